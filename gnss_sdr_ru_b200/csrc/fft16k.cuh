// fft16k.cuh -- 16000-point complex FP32 inverse DFT, one CTA, data resident in shared memory.
//
// 16000 = samples per code period at 16 MHz (SCI/*/acquisition.sci:49-50), = 20 * 20 * 40.
// Three in-register passes (Cooley-Tukey, decimation in time) with two exchanges through shared
// memory instead of a radix-2/4 ladder: shared-memory traffic, not FLOPs, bounds an on-chip FFT of
// this size (128 KB of complex data per exchange at 128 B/clk/SM), so the design minimises the number
// of exchanges: 2.  400 threads; every pass is an exact number of butterflies per thread
// (2 x DFT-20, 2 x DFT-20, 1 x DFT-40).
//
//   input index  n = 800*n1 + 40*a + b      (n1<20, a<20, b<40)
//   pass 1: DFT-20 over n1, twiddle W_16000^{(40a+b)*k1}            -> S1[k1][40a+b]
//   pass 2: DFT-20 over a,  twiddle W_800^{b*ka}                    -> S2[b][20*k1+ka]  (row stride 401: bank-conflict free)
//   pass 3: DFT-40 over b                                           -> out[k1 + 20*ka + 400*kb]
//
// The transform is the UNNORMALISED INVERSE (kernel e^{+2*pi*i*nk/N}); forward transforms are obtained
// by conjugating input and output.  Small DFTs are generated at compile time (templates, constexpr
// twiddles); inter-pass twiddles are powers of one table entry per butterfly (binary power tree,
// error ~ 5 ulp) so the only table traffic is 840 complex numbers.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

namespace fft16k {

constexpr int N = 16000, R1 = 20, R2 = 20, R3 = 40, M = R2 * R3;  // M = 800
constexpr int THREADS = 400;
constexpr int S2_STRIDE = 401;
constexpr size_t SMEM_BYTES = (size_t)R3 * S2_STRIDE * sizeof(float2);  // 128320 >= 16000*8

// ---- compile-time trigonometry ---------------------------------------------------------------
namespace ct {
constexpr double PI = 3.14159265358979323846264338327950288;
constexpr double sin_series(double x) {
  double term = x, sum = x;
  for (int i = 1; i <= 24; i++) {
    term *= -x * x / ((2.0 * i) * (2.0 * i + 1.0));
    sum += term;
  }
  return sum;
}
constexpr double cos_series(double x) {
  double term = 1.0, sum = 1.0;
  for (int i = 1; i <= 24; i++) {
    term *= -x * x / ((2.0 * i - 1.0) * (2.0 * i));
    sum += term;
  }
  return sum;
}
constexpr int wrap(int k, int n) { return ((k % n) + n) % n; }
constexpr double cosk(int k, int n) {  // cos(2*pi*k/n)
  k = wrap(k, n);
  if (k == 0) return 1.0;
  if (2 * k == n) return -1.0;
  if (4 * k == n || 4 * k == 3 * n) return 0.0;
  double x = 2.0 * PI * k / n;
  if (x > PI) x -= 2.0 * PI;
  return cos_series(x);
}
constexpr double sink(int k, int n) {  // sin(2*pi*k/n)
  k = wrap(k, n);
  if (k == 0 || 2 * k == n) return 0.0;
  if (4 * k == n) return 1.0;
  if (4 * k == 3 * n) return -1.0;
  double x = 2.0 * PI * k / n;
  if (x > PI) x -= 2.0 * PI;
  return sin_series(x);
}
}  // namespace ct

template <int I>
using IC = std::integral_constant<int, I>;
template <int B, int E, class F>
__device__ __forceinline__ void sfor(F &&f) {
  if constexpr (B < E) {
    f(IC<B>{});
    sfor<B + 1, E>(f);
  }
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// a * exp(+2*pi*i*K/Nn) with K, Nn known at compile time
template <int K, int Nn>
__device__ __forceinline__ float2 mul_tw(float2 a) {
  constexpr int k = ct::wrap(K, Nn);
  if constexpr (k == 0)
    return a;
  else if constexpr (4 * k == Nn)
    return make_float2(-a.y, a.x);
  else if constexpr (2 * k == Nn)
    return make_float2(-a.x, -a.y);
  else if constexpr (4 * k == 3 * Nn)
    return make_float2(a.y, -a.x);
  else {
    constexpr float c = (float)ct::cosk(k, Nn), s = (float)ct::sink(k, Nn);
    return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
  }
}

// ---- small inverse DFTs in registers ------------------------------------------------------------
template <int R>
struct IDft;

template <>
struct IDft<2> {
  static __device__ __forceinline__ void run(float2 (&v)[2]) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};
template <>
struct IDft<4> {
  static __device__ __forceinline__ void run(float2 (&v)[4]) {
    float2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
    float2 s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
    v[0] = cadd(s02, s13);
    v[2] = csub(s02, s13);
    v[1] = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i*d13
    v[3] = make_float2(d02.x + d13.y, d02.y - d13.x);  // d02 - i*d13
  }
};
template <>
struct IDft<5> {
  static __device__ __forceinline__ void run(float2 (&v)[5]) {
    constexpr float c1 = (float)ct::cosk(1, 5), c2 = (float)ct::cosk(2, 5);
    constexpr float s1 = (float)ct::sink(1, 5), s2 = (float)ct::sink(2, 5);
    float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
    float2 t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    float2 a1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
    float2 a2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
    float2 b1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    float2 b2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    v[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
    // inverse kernel: X1 = a1 + i*b1, X4 = a1 - i*b1, X2 = a2 + i*b2, X3 = a2 - i*b2
    v[1] = make_float2(a1.x - b1.y, a1.y + b1.x);
    v[4] = make_float2(a1.x + b1.y, a1.y - b1.x);
    v[2] = make_float2(a2.x - b2.y, a2.y + b2.x);
    v[3] = make_float2(a2.x + b2.y, a2.y - b2.x);
  }
};

// R = A*B by Cooley-Tukey in registers: n = B*n1 + n2, k = k1 + A*k2
template <int A, int B>
struct IDftComposite {
  static __device__ __forceinline__ void run(float2 (&v)[A * B]) {
    float2 u[A * B];
    sfor<0, B>([&](auto n2c) {
      constexpr int n2 = decltype(n2c)::value;
      float2 t[A];
      sfor<0, A>([&](auto n1c) {
        constexpr int n1 = decltype(n1c)::value;
        t[n1] = v[B * n1 + n2];
      });
      IDft<A>::run(t);
      sfor<0, A>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        u[k1 * B + n2] = mul_tw<n2 * k1, A * B>(t[k1]);
      });
    });
    sfor<0, A>([&](auto k1c) {
      constexpr int k1 = decltype(k1c)::value;
      float2 t[B];
      sfor<0, B>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        t[n2] = u[k1 * B + n2];
      });
      IDft<B>::run(t);
      sfor<0, B>([&](auto k2c) {
        constexpr int k2 = decltype(k2c)::value;
        v[k1 + A * k2] = t[k2];
      });
    });
  }
};
template <>
struct IDft<8> : IDftComposite<2, 4> {};
template <>
struct IDft<20> : IDftComposite<4, 5> {};
template <>
struct IDft<40> : IDftComposite<8, 5> {};

// v[k] *= w^k for k = 0..R-1 (binary power tree)
template <int R>
__device__ __forceinline__ void mul_powers(float2 (&v)[R], float2 w) {
  float2 p[R];
  p[0] = make_float2(1.f, 0.f);
  p[1] = w;
  sfor<2, R>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    p[k] = cmul(p[k / 2], p[k - k / 2]);
  });
  sfor<1, R>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    v[k] = cmul(v[k], p[k]);
  });
}

// Twiddle tables (global memory, L1 resident): tw16k[n] = exp(+2*pi*i*n/16000), n < 800;
// tw800[b] = exp(+2*pi*i*b/800), b < 40.
struct Tables {
  const float2 *tw16k;
  const float2 *tw800;
};

// Unnormalised inverse DFT of length 16000.
//   load(n)           -> float2 input element n (called for every n exactly once, coalesced in n)
//   consume(tau0, v)  -> this thread's 40 outputs: element v[kb] has output index tau0 + 400*kb
// Must be called by all 400 threads of the CTA.  sm: SMEM_BYTES of shared memory.
template <class Load, class Consume>
__device__ __forceinline__ void ifft(float2 *sm, const Tables tb, Load load, Consume consume) {
  const int tid = threadIdx.x;
  __syncthreads();  // previous user of sm is done
  // ---- pass 1 ----
#pragma unroll 1
  for (int r = 0; r < 2; r++) {
    const int n2 = tid + THREADS * r;
    float2 v[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; n1++) v[n1] = load(M * n1 + n2);
    IDft<R1>::run(v);
    mul_powers<R1>(v, __ldg(tb.tw16k + n2));
#pragma unroll
    for (int k1 = 0; k1 < R1; k1++) sm[k1 * M + n2] = v[k1];
  }
  __syncthreads();
  // ---- pass 2 ---- (reads everything before anything is overwritten)
  float2 va[R2], vb[R2];
  const int idA = tid, idB = tid + THREADS;
  const int k1A = idA / R3, bA = idA % R3, k1B = idB / R3, bB = idB % R3;
#pragma unroll
  for (int a = 0; a < R2; a++) {
    va[a] = sm[k1A * M + R3 * a + bA];
    vb[a] = sm[k1B * M + R3 * a + bB];
  }
  __syncthreads();
  IDft<R2>::run(va);
  mul_powers<R2>(va, __ldg(tb.tw800 + bA));
#pragma unroll
  for (int ka = 0; ka < R2; ka++) sm[bA * S2_STRIDE + k1A * R2 + ka] = va[ka];
  IDft<R2>::run(vb);
  mul_powers<R2>(vb, __ldg(tb.tw800 + bB));
#pragma unroll
  for (int ka = 0; ka < R2; ka++) sm[bB * S2_STRIDE + k1B * R2 + ka] = vb[ka];
  __syncthreads();
  // ---- pass 3 ----
  float2 v[R3];
#pragma unroll
  for (int b = 0; b < R3; b++) v[b] = sm[b * S2_STRIDE + tid];
  IDft<R3>::run(v);
  const int k1 = tid / R2, ka = tid % R2;
  consume(k1 + R1 * ka, v);
}

}  // namespace fft16k
