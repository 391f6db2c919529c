# stub: rnerf/ior_utils.py imports trimesh for its (unused here) Dataset class
