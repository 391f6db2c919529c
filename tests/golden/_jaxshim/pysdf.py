class SDF:  # stub: rnerf/ior_utils.py imports pysdf for its (unused here) Dataset class
    pass
