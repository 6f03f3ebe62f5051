// Development micro-benchmark: HBM write bandwidth of candidate store patterns for the branch rows
// (each thread owns R contiguous doubles of the output).  nvcc -arch=sm_100a -O3 tools/wbench.cu -o /tmp/wbench
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int V>   // R doubles per thread, V = vector width in doubles (1, 2, 4)
__global__ void per_thread_rows(double* __restrict__ out, long long nthreads_total, double a) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; t < nthreads_total; t += stride) {
    double* p = out + t * R;
    double v = a + (double)t;
    if (V == 1) {
#pragma unroll
      for (int i = 0; i < R; i++) p[i] = v + i;
    } else if (V == 2) {
#pragma unroll
      for (int i = 0; i < R; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(v + i, v - i);
    } else {
#pragma unroll
      for (int i = 0; i < R; i += 4)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "d"(v + i), "d"(v - i), "d"(v * 2), "d"(v + 1) : "memory");
    }
  }
}

// per-thread rows staged through a per-warp shared-memory transpose: every warp-wide store is 1 KB contiguous
template <int R, int CS>
__global__ void rows_warp_transposed(double* __restrict__ out, long long nthreads_total, double a) {
  __shared__ __align__(16) double sm[4][32 * R];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; t < nthreads_total; t += stride) {
    double v = a + (double)t;
    double* mine = sm[w] + lane * R;
#pragma unroll
    for (int i = 0; i < R; i += 2) *reinterpret_cast<double2*>(mine + i) = make_double2(v + i, v - i);
    __syncwarp();
    double* base = out + (t - lane) * R;
#pragma unroll
    for (int i = 0; i < R / 4; i++) {
      const int q = i * 32 + lane;
      const double2 x0 = *reinterpret_cast<const double2*>(sm[w] + 4 * q), x1 = *reinterpret_cast<const double2*>(sm[w] + 4 * q + 2);
      if (CS) asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(base + 4 * q), "d"(x0.x), "d"(x0.y), "d"(x1.x), "d"(x1.y) : "memory");
      else asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(base + 4 * q), "d"(x0.x), "d"(x0.y), "d"(x1.x), "d"(x1.y) : "memory");
    }
    __syncwarp();
  }
}

template <int R>
__global__ void per_thread_rows_cs(double* __restrict__ out, long long nthreads_total, double a) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; t < nthreads_total; t += stride) {
    double* p = out + t * R;
    double v = a + (double)t;
#pragma unroll
    for (int i = 0; i < R; i += 4)
      asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "d"(v + i), "d"(v - i), "d"(v * 2), "d"(v + 1) : "memory");
  }
}

__global__ void coalesced(double* __restrict__ out, long long n, double a) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = a + (double)i;
}
__global__ void coalesced2(double2* __restrict__ out, long long n, double a) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = make_double2(a + (double)i, a);
}

template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; r++) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const long long N = 1LL << 29;   // 4 GiB of doubles
  double* out;
  cudaMalloc(&out, N * 8 + 256);
  const int blocks = 148 * 16, tpb = 256;
  float ms;
  ms = timeit([&] { coalesced<<<blocks, tpb>>>(out, N, 1.0); });
  printf("coalesced 8B          : %7.1f GB/s\n", N * 8 / ms / 1e6);
  ms = timeit([&] { coalesced2<<<blocks, tpb>>>((double2*)out, N / 2, 1.0); });
  printf("coalesced 16B         : %7.1f GB/s\n", N * 8 / ms / 1e6);
#define RUN(R, V, OFF)                                                                             \
  ms = timeit([&] { per_thread_rows<R, V><<<blocks, 128>>>(out + OFF, N / R, 1.0); });           \
  printf("rows R=%2d vec=%d off=%d : %7.1f GB/s\n", R, V, OFF, (N / R) * R * 8 / ms / 1e6);
  RUN(24, 1, 0) RUN(24, 2, 0) RUN(24, 4, 0) RUN(36, 1, 0) RUN(36, 2, 0) RUN(36, 4, 0) RUN(6, 1, 0) RUN(6, 2, 0)
  RUN(24, 1, 1) RUN(8, 1, 0) RUN(8, 2, 0) RUN(8, 4, 0) RUN(20, 2, 0) RUN(20, 4, 0) RUN(64, 4, 0)
  ms = timeit([&] { per_thread_rows_cs<24><<<blocks, 128>>>(out, N / 24, 1.0); });
  printf("rows R=24 vec=4 .cs     : %7.1f GB/s\n", (N / 24) * 24 * 8 / ms / 1e6);
  ms = timeit([&] { per_thread_rows_cs<36><<<blocks, 128>>>(out, N / 36, 1.0); });
  printf("rows R=36 vec=4 .cs     : %7.1f GB/s\n", (N / 36) * 36 * 8 / ms / 1e6);
  ms = timeit([&] { rows_warp_transposed<24, 0><<<148 * 8, 128>>>(out, (N / 24) / 32 * 32, 1.0); });
  printf("rows R=24 warp-transposed: %7.1f GB/s\n", (N / 24) * 24 * 8 / ms / 1e6);
  ms = timeit([&] { rows_warp_transposed<36, 0><<<148 * 6, 128>>>(out, (N / 36) / 32 * 32, 1.0); });
  printf("rows R=36 warp-transposed: %7.1f GB/s\n", (N / 36) * 36 * 8 / ms / 1e6);
  ms = timeit([&] { rows_warp_transposed<36, 1><<<148 * 6, 128>>>(out, (N / 36) / 32 * 32, 1.0); });
  printf("rows R=36 warp-transposed .cs: %7.1f GB/s\n", (N / 36) * 36 * 8 / ms / 1e6);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
