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


// ---- three output arrays per thread (g: 4, J: 24, H: 36 doubles), like the branch rows of PBRARV
__device__ __forceinline__ void stv4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__global__ void rows3_direct(double* __restrict__ g, double* __restrict__ J, double* __restrict__ H, long long n, double a) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; t < n; t += stride) {
    const double v = a + (double)t;
    stv4(g + t * 4, v, v + 1, v + 2, v + 3);
#pragma unroll
    for (int i = 0; i < 24; i += 4) stv4(J + t * 24 + i, v + i, v - i, v * 2, v + 1);
#pragma unroll
    for (int i = 0; i < 36; i += 4) stv4(H + t * 36 + i, v + i, v - i, v * 3, v + 1);
  }
}
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, int bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
// per-thread padded rows in shared memory (conflict-free 16-byte stores), one bulk copy per row and array
template <int MODE>   // 0: per-thread bulk copies of padded rows; 1: linear rows, one bulk copy per warp and array
__global__ void rows3_bulk(double* __restrict__ g, double* __restrict__ J, double* __restrict__ H, long long n, double a) {
  extern __shared__ __align__(128) double sm[];
  constexpr int PG = MODE ? 4 : 6, PJ = MODE ? 24 : 26, PH = MODE ? 36 : 38;
  double* sg = sm;
  double* sJ = sg + blockDim.x * PG;
  double* sH = sJ + blockDim.x * PJ;
  const int tid = threadIdx.x, lane = tid & 31;
  long long t = blockIdx.x * (long long)blockDim.x + tid;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; t < n; t += stride) {
    const double v = a + (double)t;
    double* rg = sg + tid * PG; double* rJ = sJ + tid * PJ; double* rH = sH + tid * PH;
    *reinterpret_cast<double2*>(rg) = make_double2(v, v + 1); *reinterpret_cast<double2*>(rg + 2) = make_double2(v + 2, v + 3);
#pragma unroll
    for (int i = 0; i < 24; i += 2) *reinterpret_cast<double2*>(rJ + i) = make_double2(v + i, v - i);
#pragma unroll
    for (int i = 0; i < 36; i += 2) *reinterpret_cast<double2*>(rH + i) = make_double2(v + i, v * 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (MODE == 0) {
      bulk_store(g + t * 4, rg, 32); bulk_store(J + t * 24, rJ, 192); bulk_store(H + t * 36, rH, 288);
    } else {
      __syncwarp();
      if (lane == 0) {
        bulk_store(g + t * 4, rg, 32 * 32); bulk_store(J + t * 24, rJ, 32 * 192); bulk_store(H + t * 36, rH, 32 * 288);
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (MODE == 1) __syncwarp();
  }
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
  {
    const long long n3 = (N / 64) / 128 * 128;
    double *g3 = out, *J3 = out + n3 * 4, *H3 = J3 + n3 * 24;
    ms = timeit([&] { rows3_direct<<<blocks, 128>>>(g3, J3, H3, n3, 1.0); });
    printf("rows3 (4+24+36) direct v4        : %7.1f GB/s\n", n3 * 64 * 8 / ms / 1e6);
    cudaFuncSetAttribute(rows3_bulk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 70 * 8);
    cudaFuncSetAttribute(rows3_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 64 * 8);
    for (int bl : {148 * 2, 148 * 3, 148 * 6}) {
      ms = timeit([&] { rows3_bulk<0><<<bl, 128, 128 * 70 * 8>>>(g3, J3, H3, n3, 1.0); });
      printf("rows3 bulk per-thread rows grid=%4d: %7.1f GB/s\n", bl, n3 * 64 * 8 / ms / 1e6);
      ms = timeit([&] { rows3_bulk<1><<<bl, 128, 128 * 64 * 8>>>(g3, J3, H3, n3, 1.0); });
      printf("rows3 bulk per-warp linear  grid=%4d: %7.1f GB/s\n", bl, n3 * 64 * 8 / ms / 1e6);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
