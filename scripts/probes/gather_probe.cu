// Development aid: DRAM bytes fetched by a strided gather of 32-byte records (the access pattern of select_kernel and of
// resample_kernel's record reads) for different load flavours.  Run under ncu:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ float4 ld16(const float4* p) {
  float4 v;
  if (MODE == 0) v = __ldg(p);
  else if (MODE == 1) v = *p;
  else if (MODE == 2) v = __ldcs(p);
  else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else if (MODE == 4) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else if (MODE == 5) asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else asm volatile("ld.global.lu.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// record i of ray r at rec[(r * S + i * P) * 2 .. +1] (2 float4 = 32 B); one thread per (ray, coarse sample)
template <int MODE>
__global__ void gather(const float4* __restrict__ rec, long n_rays, int S, int P, int nc, float4* __restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n_rays * nc) return;
  const long r = i / nc; const int j = (int)(i - r * nc);
  const float4* p = rec + (r * S + (long)j * P + 5) * 2;
  const float4 a = ld16<MODE>(p), b = ld16<MODE>(p + 1);
  out[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// one 32-byte record per thread PAIR: lane 2k reads the first float4, lane 2k+1 the second -> one 32 B request per pair
template <int MODE>
__global__ void gather_pair(const float4* __restrict__ rec, long n_rays, int S, int P, int nc, float4* __restrict__ out) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long i = t >> 1; const int h = (int)(t & 1);
  if (i >= n_rays * nc) return;
  const long r = i / nc; const int j = (int)(i - r * nc);
  const float4 a = ld16<MODE>(rec + (r * S + (long)j * P + 5) * 2 + h);
  out[t] = a;
}

int main() {
  const long B = 640000; const int S = 768, P = 12, NC = 64;
  float4 *rec, *out;
  cudaMalloc(&rec, B * S * 32); cudaMalloc(&out, B * NC * 32);
  cudaMemset(rec, 0, B * S * 32);
  const long n = B * NC;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(K, M, THREADS_PER)                                                                           \
  { for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0);                                              \
      K<M><<<(unsigned)((n * THREADS_PER + 255) / 256), 256>>>(rec, B, S, P, NC, out);                    \
      cudaEventRecord(e1); cudaEventSynchronize(e1); }                                                    \
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("%s mode %d: %.3f ms  (%s)\n", #K, M, ms, cudaGetErrorString(cudaGetLastError())); }
  RUN(gather, 0, 1) RUN(gather, 1, 1) RUN(gather, 2, 1) RUN(gather, 3, 1) RUN(gather, 4, 1) RUN(gather, 5, 1) RUN(gather, 6, 1)
  RUN(gather_pair, 0, 2) RUN(gather_pair, 1, 2) RUN(gather_pair, 4, 2)
  return 0;
}
