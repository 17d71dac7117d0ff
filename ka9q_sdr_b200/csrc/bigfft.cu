// K1+K2 implementation: see bigfft.cuh.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "bigfft.cuh"
#include "fft_regs.cuh"
#include "util.cuh"
#include "bigfft_r128.cuh"

namespace k9 {

// ---------------- device side ----------------

// ring position of logical index idx of batch `batch`: base = (ring_off + batch*ring_step) mod cap is reduced once per
// thread (ring_base), then every element needs a single conditional subtract instead of a 64-bit modulo
__device__ __forceinline__ long long ring_base(const PassArgs& a, int batch) {
  return (a.ring_off + (long long)batch * a.ring_step) % a.ring_cap;
}
__device__ __forceinline__ long long ring_pos(const PassArgs& a, long long base, int idx) {
  long long pos = base + idx;  // idx < N <= cap
  return pos >= a.ring_cap ? pos - a.ring_cap : pos;
}

__device__ __forceinline__ float2 load_input(const PassArgs& a, int batch, long long base, int idx) {
  switch (a.in_mode) {
    default:
    case IN_C32:
      return reinterpret_cast<const float2*>(a.in)[(long long)batch * a.in_batch_stride + idx];
    case IN_RING_S16: {
      const long long pos = ring_pos(a, base, idx);
      short2 v = reinterpret_cast<const short2*>(a.in)[pos];
      // reference radio.c:113-114,122: (float)int16 * SCALE16, then * gain_factor
      return make_float2(((float)v.x * a.scale) * a.gain, ((float)v.y * a.scale) * a.gain);
    }
    case IN_RING_S8: {
      const long long pos = ring_pos(a, base, idx);
      char2 v = reinterpret_cast<const char2*>(a.in)[pos];
      return make_float2(((float)v.x * a.scale) * a.gain, ((float)v.y * a.scale) * a.gain);
    }
    case IN_RING_C32: {
      const long long pos = ring_pos(a, base, idx);
      return reinterpret_cast<const float2*>(a.in)[pos];
    }
    case IN_RING_R32: {
      const long long pos = ring_pos(a, base, idx);
      return make_float2(reinterpret_cast<const float*>(a.in)[pos], 0.f);
    }
  }
}

template <int SIGN>
__device__ __forceinline__ float2 big_twiddle(const PassArgs& a, unsigned e) {
  // W_N^e = lo[e & 1023] * hi[e >> 10]
  float2 lo = __ldg(a.tw_lo + (e & 1023u));
  float2 hi = __ldg(a.tw_hi + (e >> 10));
  float2 w = cmul(lo, hi);
  if (SIGN > 0) w.y = -w.y;
  return w;
}

// One Stockham pass: T columns per CTA, R = R1*R2 points per column.
// Sub-pass A: radix R1, thread (col, p) p in [0,R2): loads x[c + ncols*(p + R2*r)], r<R1; writes u[R1*p + j] * W_R^(p*j)
// Sub-pass B: radix R2, thread (col, q) q in [0,R1): loads u[q + R1*r], r<R2; output index jt = q + R1*j.
template <int R1, int R2, int T, int SIGN, bool QFAST>
__global__ void __launch_bounds__(T*(R1 > R2 ? R1 : R2)) fft_pass_kernel(const PassArgs a) {
  constexpr int R = R1 * R2;
  constexpr int S = R + 1;  // padded column stride (float2 units)
  extern __shared__ float2 sm[];
  const int tid = threadIdx.x;
  const int batch = blockIdx.y;
  const int col0 = blockIdx.x * T;
  const int s = a.N / a.n_cur;

  // ---- sub-pass A ----
  float esum = 0.f;
  const long long rbase = (a.in_mode == IN_C32) ? 0 : ring_base(a, batch);
  if (tid < T * R2) {
    const int col = tid % T;
    const int p = tid / T;
    const int c = col0 + col;
    float2 v[R1];
    if (c < a.ncols) {
#pragma unroll
      for (int r = 0; r < R1; r++) {
        const int idx = c + a.ncols * (p + R2 * r);
        v[r] = load_input(a, batch, rbase, idx);
        if (a.energy != nullptr && idx >= a.stat_from) esum += v[r].x * v[r].x + v[r].y * v[r].y;
      }
      Dft<R1, SIGN>::run(v);
#pragma unroll
      for (int j = 0; j < R1; j++) {
        float2 o = v[j];
        if (j > 0 && p > 0) {
          float2 w = __ldg(a.tw_r + p * j);
          if (SIGN > 0) w.y = -w.y;
          o = cmul(o, w);
        }
        sm[col * S + R1 * p + j] = o;
      }
    }
  }
  if (a.energy != nullptr) {
    // warp-shuffle reduction at a convergent point, one atomic per warp
    // (status only: numerator of demod->sig.if_power, reference radio.c:123,143-144)
    esum = warp_sum(esum);
    if ((tid & 31) == 0 && esum != 0.f) atomicAdd(a.energy + batch, esum);
  }
  __syncthreads();
  // ---- sub-pass B ----
  if (tid < T * R1) {
    const int col = QFAST ? tid / R1 : tid % T;
    const int q = QFAST ? tid % R1 : tid / T;
    const int c = col0 + col;
    if (c < a.ncols) {
      float2 v[R2];
#pragma unroll
      for (int r = 0; r < R2; r++) v[r] = sm[col * S + q + R1 * r];
      Dft<R2, SIGN>::run(v);
      const int qg = c % s;
      const int pg = c / s;
      float2* out = a.out + (long long)batch * a.out_batch_stride + qg + (long long)s * ((long long)R * pg);
      const bool last = (a.n_cur == R);
#pragma unroll
      for (int j = 0; j < R2; j++) {
        const int jt = q + R1 * j;
        float2 o = v[j];
        if (!last) {
          const unsigned e = (unsigned)pg * (unsigned)jt * (unsigned)s;  // < N
          if (e != 0) o = cmul(o, big_twiddle<SIGN>(a, e));
        }
        out[(long long)s * jt] = o;
      }
    }
  }
}

// ---------------- persistent variant with asynchronous staging (large forward transforms) ----------------
//
// Same pass, but the CTAs are persistent and the global->shared copies of tile i+1 are issued with cp.async (LDGSTS) while
// tile i is transformed, into a per-thread staging area (each thread later reads back exactly the slots it copied, so the
// staging needs no barrier). ncu on the synchronous kernel showed 7-15 warps per issue slot parked on long_scoreboard and
// ~1.8 TB/s; this keeps loads in flight during the butterflies without spending registers on them.
__device__ __forceinline__ void cp_async_bytes(void* smem_dst, const void* gsrc, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int R1, int R2, int T, bool QFAST>
__global__ void __launch_bounds__(T*(R1 > R2 ? R1 : R2)) fft_pass_async_kernel(const PassArgs a) {
  constexpr int SIGN = -1;
  constexpr int R = R1 * R2;
  constexpr int S = R + 1;     // padded column stride of the exchange buffer (float2 units)
  constexpr int NA = T * R2;   // threads active in sub-pass A
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* sm = reinterpret_cast<float2*>(smraw);                                       // exchange buffer [T*S]
  unsigned long long* stage = reinterpret_cast<unsigned long long*>(sm + T * S + 1);  // [2][R1][NA] 8-byte slots
  const int tid = threadIdx.x;
  const int batch = blockIdx.y;
  const int s = a.N / a.n_cur;
  const int ntiles = (a.ncols + T - 1) / T;
  const bool actA = tid < NA;
  const int colA = tid % T;
  const int pA = tid / T;
  const bool is_c32 = (a.in_mode == IN_C32) || (a.in_mode == IN_RING_C32);
  const int esz = is_c32 ? 8 : 4;  // IN_RING_S16: short2; IN_RING_R32: float
  const long long rbase = (a.in_mode == IN_C32) ? 0 : ring_base(a, batch);
  const char* gbase = reinterpret_cast<const char*>(a.in) +
                      (a.in_mode == IN_C32 ? (long long)batch * a.in_batch_stride * 8 : 0);

  auto issue = [&](int tile, int buf) {
    const int c = tile * T + colA;
    if (actA && c < a.ncols) {
#pragma unroll
      for (int r = 0; r < R1; r++) {
        const int idx = c + a.ncols * (pA + R2 * r);
        const long long pos = (a.in_mode == IN_C32) ? idx : ring_pos(a, rbase, idx);
        cp_async_bytes(&stage[(buf * R1 + r) * NA + tid], gbase + pos * esz, esz);
      }
    }
    cp_async_commit();
  };

  int tile = blockIdx.x;
  int buf = 0;
  if (tile < ntiles) issue(tile, 0);
  for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < ntiles)
      issue(next, buf ^ 1);
    else
      cp_async_commit();  // empty group: keeps "all but the newest group" == "this tile" for wait_group<1>
    cp_async_wait<1>();
    const int col0 = tile * T;
    // ---- sub-pass A ----
    float esum = 0.f;
    if (actA) {
      const int c = col0 + colA;
      if (c < a.ncols) {
        float2 v[R1];
#pragma unroll
        for (int r = 0; r < R1; r++) {
          const unsigned long long raw = stage[(buf * R1 + r) * NA + tid];
          float2 x;
          if (is_c32) {
            x = make_float2(__uint_as_float((unsigned)raw), __uint_as_float((unsigned)(raw >> 32)));
          } else if (a.in_mode == IN_RING_S16) {
            const short lo = (short)(raw & 0xffffu), hi = (short)((raw >> 16) & 0xffffu);
            // reference radio.c:113-114,122: (float)int16 * SCALE16, then * gain_factor
            x = make_float2(((float)lo * a.scale) * a.gain, ((float)hi * a.scale) * a.gain);
          } else {
            x = make_float2(__uint_as_float((unsigned)raw), 0.f);
          }
          v[r] = x;
          if (a.energy != nullptr && c + a.ncols * (pA + R2 * r) >= a.stat_from) esum += x.x * x.x + x.y * x.y;
        }
        Dft<R1, SIGN>::run(v);
#pragma unroll
        for (int j = 0; j < R1; j++) {
          float2 o = v[j];
          if (j > 0) o = cmul(o, __ldg(a.tw_r + pA * j));  // tw_r[0] == 1 exactly: no test on pA
          sm[colA * S + R1 * pA + j] = o;
        }
      }
    }
    if (a.energy != nullptr) {
      esum = warp_sum(esum);
      if ((tid & 31) == 0 && esum != 0.f) atomicAdd(a.energy + batch, esum);
    }
    __syncthreads();
    // ---- sub-pass B ----
    if (tid < T * R1) {
      const int col = QFAST ? tid / R1 : tid % T;
      const int q = QFAST ? tid % R1 : tid / T;
      const int c = col0 + col;
      if (c < a.ncols) {
        float2 w[R2];
#pragma unroll
        for (int r = 0; r < R2; r++) w[r] = sm[col * S + q + R1 * r];
        Dft<R2, SIGN>::run(w);
        const int qg = c % s;
        const int pg = c / s;
        float2* out = a.out + (long long)batch * a.out_batch_stride + qg + (long long)s * ((long long)R * pg);
        if (a.n_cur != R) {
          // inter-pass twiddles W_N^(pg*jt*s): all table loads are issued before the first product (no per-output
          // branch: entry 0 of both tables is exactly 1, so e == 0 needs no special case)
          const unsigned es = (unsigned)pg * (unsigned)s;
          float2 lo[R2], hi[R2];
#pragma unroll
          for (int j = 0; j < R2; j++) {
            const unsigned e = es * (unsigned)(q + R1 * j);  // < N
            lo[j] = __ldg(a.tw_lo + (e & 1023u));
            hi[j] = __ldg(a.tw_hi + (e >> 10));
          }
#pragma unroll
          for (int j = 0; j < R2; j++) w[j] = cmul(w[j], cmul(lo[j], hi[j]));
        }
#pragma unroll
        for (int j = 0; j < R2; j++) out[(long long)s * (q + R1 * j)] = w[j];
      }
    }
    __syncthreads();  // the exchange buffer is reused by the next tile
  }
  cp_async_wait<0>();
}

template <int R1, int R2, int T>
static cudaError_t launch_async_kind(const PassArgs& a, int batch, bool qfast, cudaStream_t st) {
  constexpr int R = R1 * R2;
  constexpr int threads = T * (R1 > R2 ? R1 : R2);
  const size_t smem = sizeof(float2) * ((size_t)T * (R + 1) + 1) + 8ull * 2 * R1 * T * R2;
  const int ntiles = (a.ncols + T - 1) / T;
  static int ctas_per_sm = -1;
  if (ctas_per_sm < 0) {
    const char* e = getenv("KA9Q_B200_FFT_CTAS");
    ctas_per_sm = e ? atoi(e) : 4;
  }
  int gx = (148 * ctas_per_sm + batch - 1) / batch;  // persistent CTAs per SM over the whole batch
  if (gx > ntiles) gx = ntiles;
  dim3 grid(gx, batch);
  if (qfast) {
    cudaFuncSetAttribute(fft_pass_async_kernel<R1, R2, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fft_pass_async_kernel<R1, R2, T, true><<<grid, threads, smem, st>>>(a);
  } else {
    cudaFuncSetAttribute(fft_pass_async_kernel<R1, R2, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fft_pass_async_kernel<R1, R2, T, false><<<grid, threads, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// returns cudaErrorNotSupported when this pass should use the synchronous kernel
static cudaError_t launch_pass_async(int R1, int R2, const PassArgs& a, int batch, int sign, bool qfast, cudaStream_t st) {
  if (sign >= 0 || a.in_mode == IN_RING_S8) return cudaErrorNotSupported;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("KA9Q_B200_FFT_ASYNC");
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled) return cudaErrorNotSupported;
#define K9_ACASE(r1, r2, t)                                                                   \
  if (R1 == r1 && R2 == r2) {                                                                 \
    if ((long long)((a.ncols + t - 1) / t) * batch < 148 * 3 * 2) return cudaErrorNotSupported; \
    return launch_async_kind<r1, r2, t>(a, batch, qfast, st);                                 \
  }
  K9_ACASE(8, 16, 16)
  K9_ACASE(10, 16, 16)
  K9_ACASE(5, 16, 16)
  K9_ACASE(8, 8, 32)
#undef K9_ACASE
  return cudaErrorNotSupported;
}

// ---------------- host side ----------------

struct PassKind {
  int R1, R2, T;
};
static const PassKind kKinds[] = {
    {16, 16, 16}, {20, 16, 16}, {10, 16, 16}, {8, 16, 16}, {5, 16, 16}, {8, 8, 32}, {4, 8, 32},
    {4, 4, 64},   {20, 20, 8},  {10, 10, 16}, {5, 10, 16}, {5, 5, 32},  {3, 16, 16}, {6, 16, 16}, {12, 16, 16}, {15, 16, 16},
};
static const int kNumKinds = sizeof(kKinds) / sizeof(kKinds[0]);

template <int R1, int R2, int T>
static cudaError_t launch_kind(const PassArgs& a, int batch, int sign, bool qfast, cudaStream_t st) {
  constexpr int R = R1 * R2;
  constexpr int threads = T * (R1 > R2 ? R1 : R2);
  const size_t smem = sizeof(float2) * (size_t)T * (R + 1);
  dim3 grid((a.ncols + T - 1) / T, batch);
  static_assert(sizeof(float2) * (size_t)T * (R + 1) <= 48 * 1024, "pass tile must fit the default smem window");
  auto set = [](auto) {};
  if (sign < 0) {
    if (qfast) {
      set(fft_pass_kernel<R1, R2, T, -1, true>);
      fft_pass_kernel<R1, R2, T, -1, true><<<grid, threads, smem, st>>>(a);
    } else {
      set(fft_pass_kernel<R1, R2, T, -1, false>);
      fft_pass_kernel<R1, R2, T, -1, false><<<grid, threads, smem, st>>>(a);
    }
  } else {
    if (qfast) {
      set(fft_pass_kernel<R1, R2, T, +1, true>);
      fft_pass_kernel<R1, R2, T, +1, true><<<grid, threads, smem, st>>>(a);
    } else {
      set(fft_pass_kernel<R1, R2, T, +1, false>);
      fft_pass_kernel<R1, R2, T, +1, false><<<grid, threads, smem, st>>>(a);
    }
  }
  return cudaGetLastError();
}

static cudaError_t launch_pass(int R1, int R2, const PassArgs& a, int batch, int sign, bool qfast, cudaStream_t st) {
#define K9_CASE(r1, r2, t) \
  if (R1 == r1 && R2 == r2) return launch_kind<r1, r2, t>(a, batch, sign, qfast, st);
  K9_CASE(16, 16, 16)
  K9_CASE(20, 16, 16)
  K9_CASE(10, 16, 16)
  K9_CASE(8, 16, 16)
  K9_CASE(5, 16, 16)
  K9_CASE(8, 8, 32)
  K9_CASE(4, 8, 32)
  K9_CASE(4, 4, 64)
  K9_CASE(20, 20, 8)
  K9_CASE(10, 10, 16)
  K9_CASE(5, 10, 16)
  K9_CASE(5, 5, 32)
  K9_CASE(3, 16, 16)
  K9_CASE(6, 16, 16)
  K9_CASE(12, 16, 16)
  K9_CASE(15, 16, 16)
#undef K9_CASE
  return cudaErrorInvalidValue;
}

// Enumerate factorisations into supported pass sizes; prefer the fewest passes, then the smallest largest pass
// (small tiles = more CTAs in flight), then the largest smallest pass.
struct FactorBest {
  int depth = 99;
  int maxR = 1 << 30;
  int minR = 0;
  int pick[4];
};
static void factor_dfs(long long n, int depth, int start, int* pick, FactorBest& best) {
  if (n == 1) {
    if (depth == 0) return;
    int mx = 0, mn = 1 << 30;
    for (int i = 0; i < depth; i++) {
      int R = kKinds[pick[i]].R1 * kKinds[pick[i]].R2;
      mx = R > mx ? R : mx;
      mn = R < mn ? R : mn;
    }
    bool better = depth < best.depth || (depth == best.depth && (mx < best.maxR || (mx == best.maxR && mn > best.minR)));
    if (better) {
      best.depth = depth;
      best.maxR = mx;
      best.minR = mn;
      for (int i = 0; i < depth; i++) best.pick[i] = pick[i];
    }
    return;
  }
  if (depth == 4 || depth >= best.depth) return;
  for (int i = start; i < kNumKinds; i++) {
    const int R = kKinds[i].R1 * kKinds[i].R2;
    if (n % R == 0) {
      pick[depth] = i;
      factor_dfs(n / R, depth + 1, i, pick, best);
    }
  }
}

int bigfft_factorize(int N, int* R1, int* R2) {
  if (N < 16) return -1;
  FactorBest best;
  int pick[4];
  // allow equal-depth alternatives to be compared: search with depth bound relaxed by re-running per depth
  for (int maxd = 1; maxd <= 4; maxd++) {
    best = FactorBest();
    best.depth = maxd + 1;
    factor_dfs(N, 0, 0, pick, best);
    if (best.depth <= maxd) break;
  }
  if (best.depth > 4) return -1;
  const int d = best.depth;
  // ascending size: the strided first pass is the small one, the contiguous last pass the large one
  for (int i = 0; i < d; i++)
    for (int j = i + 1; j < d; j++) {
      int ri = kKinds[best.pick[i]].R1 * kKinds[best.pick[i]].R2, rj = kKinds[best.pick[j]].R1 * kKinds[best.pick[j]].R2;
      if (rj < ri) {
        int t = best.pick[i];
        best.pick[i] = best.pick[j];
        best.pick[j] = t;
      }
    }
  for (int i = 0; i < d; i++) {
    R1[i] = kKinds[best.pick[i]].R1;
    R2[i] = kKinds[best.pick[i]].R2;
  }
  return d;
}

int bigfft_plan_create(BigFftPlan* plan, int N) {
  memset(plan, 0, sizeof(*plan));
  int np = bigfft_factorize(N, plan->R1, plan->R2);
  if (np < 0) return -1;
  plan->N = N;
  plan->npass = np;
  cudaGetDevice(&plan->device);
  // W_N tables in double, rounded once to float
  const int nlo = 1024;
  const int nhi = (N + 1023) / 1024;
  std::vector<float2> lo(nlo), hi(nhi);
  for (int a = 0; a < nlo; a++) {
    double ang = -2.0 * M_PI * (double)(a % N) / (double)N;
    lo[a] = make_float2((float)cos(ang), (float)sin(ang));
  }
  for (int b = 0; b < nhi; b++) {
    double ang = -2.0 * M_PI * (double)(((long long)b * 1024) % N) / (double)N;
    hi[b] = make_float2((float)cos(ang), (float)sin(ang));
  }
  if (cudaMalloc(&plan->tw_lo, sizeof(float2) * nlo) != cudaSuccess) return -2;
  if (cudaMalloc(&plan->tw_hi, sizeof(float2) * nhi) != cudaSuccess) return -2;
  cudaMemcpy(plan->tw_lo, lo.data(), sizeof(float2) * nlo, cudaMemcpyHostToDevice);
  cudaMemcpy(plan->tw_hi, hi.data(), sizeof(float2) * nhi, cudaMemcpyHostToDevice);
  for (int p = 0; p < np; p++) {
    const int R = plan->R1[p] * plan->R2[p];
    std::vector<float2> tr(R);
    for (int a = 0; a < R; a++) {
      double ang = -2.0 * M_PI * (double)a / (double)R;
      tr[a] = make_float2((float)cos(ang), (float)sin(ang));
    }
    if (cudaMalloc(&plan->tw_r[p], sizeof(float2) * R) != cudaSuccess) return -2;
    cudaMemcpy(plan->tw_r[p], tr.data(), sizeof(float2) * R, cudaMemcpyHostToDevice);
  }
  return 0;
}

void bigfft_plan_destroy(BigFftPlan* plan) {
  if (!plan) return;
  cudaFree(plan->tw_lo);
  cudaFree(plan->tw_hi);
  for (int p = 0; p < 4; p++) cudaFree(plan->tw_r[p]);
  memset(plan, 0, sizeof(*plan));
}

void bigfft_set_carveout(int pct) { set_pass128_carveout(pct); }

bool bigfft_can_route(const BigFftPlan* plan) {
  static const bool use128 = !(getenv("KA9Q_B200_FFT_R128") && atoi(getenv("KA9Q_B200_FFT_R128")) == 0);
  const int p = plan->npass - 1;
  return use128 && plan->npass >= 2 && plan->R1[p] == 10 && plan->R2[p] == 16 && (plan->N / 160) % 16 == 0 &&
         (long long)plan->N < (1ll << 30);
}

int bigfft_exec(const BigFftPlan* plan, const BigFftIn& in, float2* out, long long out_batch_stride, float2* tmp0,
                float2* tmp1, int batch, int sign, cudaStream_t stream) {
  const int N = plan->N;
  int n_cur = N;
  const void* cur_in = in.in;
  long long cur_stride = in.in_batch_stride;
  int cur_mode = in.in_mode;
  for (int p = 0; p < plan->npass; p++) {
    const int R = plan->R1[p] * plan->R2[p];
    PassArgs a;
    memset(&a, 0, sizeof(a));
    a.in = cur_in;
    a.in_batch_stride = cur_stride;
    a.in_mode = cur_mode;
    const bool lastpass = (p == plan->npass - 1);
    float2* dst;
    long long dst_stride;
    if (lastpass) {
      dst = out;
      dst_stride = out_batch_stride;
    } else {
      // ping-pong: never write the buffer being read
      dst = (cur_in == (const void*)tmp0) ? tmp1 : tmp0;
      dst_stride = N;
      if (!dst) return -3;
    }
    a.out = dst;
    a.out_batch_stride = dst_stride;
    a.N = N;
    a.n_cur = n_cur;
    a.ncols = N / R;
    a.ring_cap = in.ring_cap;
    a.ring_off = in.ring_off;
    a.ring_step = in.ring_step;
    a.scale = in.scale;
    a.gain = in.gain;
    a.stat_from = in.stat_from;
    a.energy = (p == 0) ? in.energy : nullptr;
    a.tw_lo = plan->tw_lo;
    a.tw_hi = plan->tw_hi;
    a.tw_r = plan->tw_r[p];
    if (lastpass && in.route_mask && bigfft_can_route(plan)) {
      a.route_mask = in.route_mask;
      for (int r = 0; r < 16; r++) a.route_delta[r] = in.route_delta[r];
    }
    static const bool use128 = !(getenv("KA9Q_B200_FFT_R128") && atoi(getenv("KA9Q_B200_FFT_R128")) == 0);
    cudaError_t e = use128 ? launch_pass128(plan->R1[p], plan->R2[p], a, batch, sign, stream) : cudaErrorNotSupported;
    if (e == cudaErrorNotSupported)
      e = launch_pass_async(plan->R1[p], plan->R2[p], a, batch, sign, /*qfast=*/(N / n_cur) == 1, stream);
    if (e == cudaErrorNotSupported)
      e = launch_pass(plan->R1[p], plan->R2[p], a, batch, sign, /*qfast=*/(N / n_cur) == 1, stream);
    if (e != cudaSuccess) {
      fprintf(stderr, "bigfft: launch failed: %s\n", cudaGetErrorString(e));
      return -4;
    }
    n_cur /= R;
    cur_in = dst;
    cur_stride = dst_stride;
    cur_mode = IN_C32;
  }
  return 0;
}

}  // namespace k9
