// gpssdr_acq.cu -- the GPS-SDR fixed-point FFT acquisition on the GPU, bit exact.
//
// What it replaces (RT = trunk/GNSS_SOFTWARE_RECEIVERS/REALTIME_RECEIVERS/GPS/
// GPS_SDR_REAL_TIME_GPS_RECEIVER of the reference):
//   Acquisition::doPrepIF      RT/objects/acquisition.cpp:182-236   250/500/750 Hz offsets, mix to baseband,
//                                                                   forward FFT of every millisecond, padded rows
//   Acquisition::doAcqStrong   RT/objects/acquisition.cpp:244-302   1 ms
//   Acquisition::doAcqMedium   RT/objects/acquisition.cpp:309-425   10 ms coherent, 25 Hz post-correlation DFT; reads
//                                                                   rows lcv2*20 + lcv3 (see gnssb200.h on the rows
//                                                                   a 10-ms preparation does not fill)
//   Acquisition::doAcqWeak     RT/objects/acquisition.cpp:433-570   10 ms coherent x 15 non-coherent, 25 Hz
//                                                                   post-correlation DFT, code-Doppler shift,
//                                                                   even / odd 10-ms alignment
// in the arithmetic of the reference's portable primitives: the 2048-point int16 radix-2 FFT with per-rank
// scaling flags and (x+8192)>>14 rounding (RT/objects/fft.cpp:173-233,314-440), x86_cmulsc / x86_cacc /
// x86_cmag / x86_max (RT/simd/x86.cpp:181-294).  Every int16 store wraps exactly like the C code's.
//
// Mapping: the whole search is embarrassingly parallel over (satellite, kHz bin, 250 Hz offset, even/odd):
// one CTA per combination keeps the ten 2048-point coherent rows (80 KB) and the 10 x 2048 power matrix
// (80 KB) in shared memory for all 15 non-coherent rounds; only the padded spectra rows (L2 resident) are
// read and one (maximum, index) pair per combination is written.  The final pick over combinations follows
// the reference's loop order on the host.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int NS = 2048, ROWLEN = NS + 201, NT = 1024, SCR = NS + NS / 32;  // SCR: padded scratch of fft2048

struct Cpx16 {
  int16_t i, q;
};
__device__ __forceinline__ Cpx16 unpack16(uint32_t v) {
  Cpx16 c;
  c.i = (int16_t)(v & 0xffffu);
  c.q = (int16_t)(v >> 16);
  return c;
}
__device__ __forceinline__ uint32_t pack16(int i, int q) { return ((uint32_t)i & 0xffffu) | ((uint32_t)q << 16); }

// x86_cmulsc (x86.cpp:181-216): (a*b + round) >> shift, truncated to int16
__device__ __forceinline__ uint32_t cmul_shift(uint32_t a, uint32_t b, int shift) {
  const Cpx16 A = unpack16(a), B = unpack16(b);
  int ti = (int)A.i * B.i - (int)A.q * B.q, tq = (int)A.i * B.q + (int)A.q * B.i;
  const int round = 1 << (shift - 1);
  ti = (ti + round) >> shift;
  tq = (tq + round) >> shift;
  return pack16(ti, tq);
}

// 2048-point FFT of the reference (fft.cpp): bit-reversal shuffle (doShuffle :298-310), then 11 ranks of 1024
// butterflies (bfly / bfly_noscale :401-440).  in: the 2048 packed CPX in natural order, PADDED by one word per 32
// (element j at in[PAD(j)]) so that the bit-reversed gather below -- 32 consecutive targets differ only in the
// high source-address bits -- does not land a whole warp on one bank.  x: result, natural order.  tw: the 1024
// twiddles (i, q) of W or iW.  RF bit r = scale rank r by >>1 first (a compile-time pattern: R1 = none for the
// forward transform, R2 = ranks 7 and 9 for the inverse one).  All NT threads of the CTA.
// Ranks 0-4 only combine elements inside aligned groups of 32: they run in registers, one element per lane; the
// upper lane of a pair forms the product B*W, the lower lane keeps A, one shuffle swaps the two.  Ranks 5-10 go
// through shared memory (bank-conflict free from there on).
// Arithmetic note: the reference stores the rounded product in an int16 before adding it to A; A +- product is
// stored as int16 again, so only the low 16 bits of the sum matter and the intermediate truncation can be skipped
// wherever the product stays in a 32-bit register.
#define PAD(j) ((j) + ((j) >> 5))
constexpr unsigned RF_NONE = 0u, RF_R2 = (1u << 7) | (1u << 9);  // R1 / R2 of acquisition.cpp:75-76
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)v >> 16; }
template <unsigned RF>
__device__ void fft2048(const uint32_t *in, uint32_t *x, const uint32_t *__restrict__ tw) {
  const int tid = threadIdx.x, lane = tid & 31;
  for (int l = tid; l < NS; l += NT) {  // a warp owns the aligned group of 32 around l
    const int src = (int)(__brev((unsigned)l) >> 21);
    uint32_t v = in[PAD(src)];
#pragma unroll
    for (int r = 0; r < 5; r++) {
      const int bsize = 1 << r;
      const bool upper = (lane >> r) & 1;
      const uint32_t w = __ldg(tw + ((lane & (bsize - 1)) << (10 - r)));
      int vi = lo16(v), vq = hi16(v);
      if ((RF >> r) & 1) {
        vi >>= 1;
        vq >>= 1;
      }
      const int wi = lo16(w), wq = hi16(w);
      const int pi = (vi * wi + 8192 - vq * wq) >> 14, pq = (vi * wq + 8192 + vq * wi) >> 14;  // used by the upper lanes
      const uint32_t got = __shfl_xor_sync(0xffffffffu, upper ? pack16(pi, pq) : pack16(vi, vq), bsize);
      const int gi = lo16(got), gq = hi16(got);
      // lower lane: A + P with A its own element and P received; upper lane: A - P with A received and P its own
      v = pack16(gi + (upper ? -pi : vi), gq + (upper ? -pq : vq));
    }
    x[l] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 5; r < 11; r++) {
    const int bsize = 1 << r, nblocks = 1024 >> r;
    for (int t = tid; t < 1024; t += NT) {
      const int blk = t >> r, j = t & (bsize - 1);
      const int ia = (blk << (r + 1)) + j, ib = ia + bsize;
      const uint32_t a = x[ia], b = x[ib], w = __ldg(tw + j * nblocks);
      int ai = lo16(a), aq = hi16(a), bi = lo16(b), bq = hi16(b);
      if ((RF >> r) & 1) {
        ai >>= 1; aq >>= 1; bi >>= 1; bq >>= 1;
      }
      const int wi = lo16(w), wq = hi16(w);
      const int pi = (bi * wi + 8192 - bq * wq) >> 14, pq = (bi * wq + 8192 + bq * wi) >> 14;
      x[ib] = pack16(ai - pi, aq - pq);
      x[ia] = pack16(ai + pi, aq + pq);
    }
    __syncthreads();
  }
}

// doPrepIF: one CTA per (offset, millisecond) row
__global__ void __launch_bounds__(NT) gsa_prep_kernel(const uint32_t *iq, int ms, const uint32_t *wipe /* [4][10*2048] */,
                                                      const uint32_t *twf, uint32_t *rows /* [4*ms][ROWLEN] */) {
  __shared__ uint32_t x[NS], scratch[SCR];
  const int row = blockIdx.x, off = row / ms, m = row % ms, tid = threadIdx.x;
  for (int j = tid; j < NS; j += NT)
    scratch[PAD(j)] = cmul_shift(iq[(size_t)m * NS + j], wipe[(size_t)off * 10 * NS + (m % 10) * NS + j], 14);
  __syncthreads();
  fft2048<RF_NONE>(scratch, x, twf);  // R1: no rank is scaled
  uint32_t *p = rows + (size_t)row * ROWLEN;
  for (int j = tid; j < ROWLEN - 1; j += NT) p[j] = x[(j + NS - 100) & (NS - 1)];  // 100 wrapped bins on either side
  if (tid == 0) p[ROWLEN - 1] = 0;
}

// dft_rows[r][j] of the post-correlation DFT as (i, nq, q, ni): constant-bank operands of the multiply-adds
__constant__ int4 c_dft[100];

struct Best {
  int mag, idx;
};
// first maximum strictly greater than zero, lowest index among equals (x86_max :273-294)
__device__ Best block_first_max(int mag, int idx, Best *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const int m2 = __shfl_xor_sync(0xffffffffu, mag, o), i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    if (m2 > mag || (m2 == mag && i2 < idx)) {
      mag = m2;
      idx = i2;
    }
  }
  if (lane == 0) red[warp] = Best{mag, idx};
  __syncthreads();
  Best b = red[0];
  for (int w = 1; w < NT / 32; w++)
    if (red[w].mag > b.mag || (red[w].mag == b.mag && red[w].idx < b.idx)) b = red[w];
  __syncthreads();
  return b;
}

// doAcqStrong: one CTA per (sv, kHz bin, 250 Hz offset)
__global__ void __launch_bounds__(NT) gsa_strong_kernel(const uint32_t *rows, const uint32_t *codes, const int *sv_list, int nd, int l0,
                                                        const uint32_t *twi, Best *out) {
  __shared__ uint32_t x[NS], scratch[SCR];
  __shared__ Best red[NT / 32];
  const int combo = blockIdx.x % (nd * 4), svi = blockIdx.x / (nd * 4);
  const int l = l0 + combo / 4, l2 = combo % 4, tid = threadIdx.x;
  const uint32_t *code = codes + (size_t)sv_list[svi] * NS;
  const uint32_t *row = rows + (size_t)l2 * ROWLEN + 100 + l;
  for (int j = tid; j < NS; j += NT) scratch[PAD(j)] = cmul_shift(row[j], code[j], 10);
  __syncthreads();
  fft2048<RF_R2>(scratch, x, twi);
  int mag = 0, idx = 0;
  for (int j = tid; j < NS; j += NT) {
    const Cpx16 c = unpack16(x[j]);
    const int p = (int)c.i * c.i + (int)c.q * c.q;  // x86_cmag
    if (p > mag) {
      mag = p;
      idx = j;
    }
  }
  const Best b = block_first_max(mag, idx, red);
  if (tid == 0) out[blockIdx.x] = b;
}

// doAcqWeak: one CTA per (sv, kHz bin, 250 Hz offset, even/odd); doAcqMedium (MEDIUM): one CTA per (sv, kHz bin,
// lcv2), a single round over rows lcv2*20 + lcv3 with the multiply shifted by 10 and no code-Doppler shift
template <bool MEDIUM>
__global__ void __launch_bounds__(NT) gsa_weak_kernel(const uint32_t *rows, const uint32_t *codes, const int *sv_list, int nd, int l0,
                                                      const uint32_t *twi, Best *out) {
  extern __shared__ uint32_t sm[];
  uint32_t *coh = sm;                       // [10][2048] packed CPX
  int *power = (int *)(sm + 10 * NS);       // [10][2048]
  uint32_t *scratch = sm + 20 * NS;         // [SCR]
  __shared__ Best red[NT / 32];
  const int per_sv = nd * (MEDIUM ? 4 : 8), combo = blockIdx.x % per_sv, svi = blockIdx.x / per_sv;
  const int l = l0 + combo / (MEDIUM ? 4 : 8), l2 = MEDIUM ? combo % 4 : (combo / 2) % 4, k = MEDIUM ? 0 : combo % 2, tid = threadIdx.x;
  const uint32_t *code = codes + (size_t)sv_list[svi] * NS;
  for (int j = tid; j < 10 * NS; j += NT) power[j] = 0;
  __syncthreads();
  for (int i = 0; i < (MEDIUM ? 1 : 15); i++) {
    for (int l3 = 0; l3 < 10; l3++) {
      const int r = MEDIUM ? l2 * 20 + l3 : l2 * 310 + l3 + i * 20 + k * 10;
      const uint32_t *row = rows + (size_t)r * ROWLEN + 100 + l;
      uint32_t *x = coh + l3 * NS;
      for (int j = tid; j < NS; j += NT) scratch[PAD(j)] = cmul_shift(row[j], code[j], MEDIUM ? 10 : 9);
      __syncthreads();
      fft2048<RF_R2>(scratch, x, twi);
    }
    // code-Doppler shift of this round (acquisition.cpp:486-492)
    const double doppler = (double)(l * 1000) + (float)(l2 * 250);
    const double code_doppler = (double)i * .02 * 2048000 * doppler / 1.57542e9;
    const int shift = (int)floor(code_doppler);
    for (int d = tid; d < NS; d += NT) {
      int di[10], dq[10];
#pragma unroll
      for (int j = 0; j < 10; j++) {
        const Cpx16 c = unpack16(coh[j * NS + d]);
        di[j] = c.i;
        dq[j] = c.q;
      }
      const int col = (d + shift + NS) % NS;
#pragma unroll
      for (int r = 0; r < 10; r++) {
        int ia = 0, qa = 0;  // x86_cacc :220-251
#pragma unroll
        for (int j = 0; j < 10; j++) {
          const int4 w = c_dft[r * 10 + j];
          ia += di[j] * w.x + dq[j] * w.y;
          qa += di[j] * w.z + dq[j] * w.w;
        }
        const int ti = (int16_t)(ia >> 16), tq = (int16_t)(qa >> 16);
        power[r * NS + col] += ti * ti + tq * tq;
      }
    }
    __syncthreads();
  }
  int mag = 0, idx = 0;
  for (int j = tid; j < 10 * NS; j += NT)
    if (power[j] > mag) {
      mag = power[j];
      idx = j;
    }
  const Best b = block_first_max(mag, idx, red);
  if (tid == 0) out[blockIdx.x] = b;
}

// ---- host-side tables, same arithmetic as the reference's generators (run on the host's libm) ----
void make_twiddles(std::vector<uint32_t> &fwd, std::vector<uint32_t> &inv) {  // FFT::initW, fft.cpp:121-149
  const double pi = 3.14159265358979323846264338327;
  fwd.resize(1024);
  inv.resize(1024);
  for (int l = 0; l < 1024; l++) {
    const double phase = (-2 * pi * l) / 2048;
    const short c = (short)floor(16384 * cos(phase)), s = (short)floor(16384 * sin(phase));
    fwd[l] = ((uint32_t)(uint16_t)c) | ((uint32_t)(uint16_t)s << 16);
    inv[l] = ((uint32_t)(uint16_t)c) | ((uint32_t)(uint16_t)(short)(-s) << 16);
  }
}
void make_wipeoff(double fif, std::vector<uint32_t> &w) {  // sine_gen, misc.cpp:95-114 (float phase), acquisition.cpp:112-118
  w.resize((size_t)4 * 10 * NS);
  for (int k = 0; k < 4; k++) {
    const double f = -fif - 250.0 * k, fs = 2048000;
    float phase = 0, phase_step = (float)6.283185307179586 * f / fs;
    for (int l = 0; l < 10 * NS; l++) {
      const int16_t c = (int16_t)floor(16383.0 * cosf(phase)), s = (int16_t)floor(16383.0 * sinf(phase));
      w[(size_t)k * 10 * NS + l] = ((uint32_t)(uint16_t)c) | ((uint32_t)(uint16_t)s << 16);
      phase += phase_step;
    }
  }
}
void make_dft(std::vector<int2> &d) {  // wipeoff_gen, misc.cpp:148-166; acquisition.cpp:107-109
  d.resize(100);
  for (int r = 0; r < 10; r++) {
    double phase = 0;
    const double phase_step = (double)6.283185307179586 * ((float)r * 25.0 - 112.5) / 1000.0;
    for (int l = 0; l < 10; l++) {
      const int16_t c = (int16_t)floor(16383.0 * cos(phase)), s = (int16_t)floor(16383.0 * sin(phase));
      const int16_t ns = (int16_t)-s;
      d[r * 10 + l].x = (int)(((uint32_t)(uint16_t)c) | ((uint32_t)(uint16_t)ns << 16));  // i, nq
      d[r * 10 + l].y = (int)(((uint32_t)(uint16_t)s) | ((uint32_t)(uint16_t)c << 16));   // q, ni
      phase += phase_step;
    }
  }
}

}  // namespace

static int prep_ms(int type) { return type == 0 ? 1 : (type == 1 ? 10 : (type == 2 ? 310 : 0)); }

// prior_iq / prior_type: an earlier doPrepIF of the same object whose rows persist under this one (medium only)
static int gpssdr_run(gnssb200_handle *h, const int16_t *iq, int type, const int16_t *prior_iq, int prior_type, double fif,
                      const int16_t *prn_codes, int n_codes, const int32_t *sv_list, int n_sv, int doppmin, int doppmax,
                      gnssb200_gpssdr_result *results) {
  const int ms = prep_ms(type), pms = prior_iq ? prep_ms(prior_type) : 0;
  // doAcqStrong / doAcqWeak stop before doppmax/1000, doAcqMedium includes it (acquisition.cpp:258,325,450)
  const int l0 = doppmin / 1000, nd = doppmax / 1000 - doppmin / 1000 + (type == 1 ? 1 : 0);
  if (!h || !iq || !prn_codes || !sv_list || !results || ms == 0 || (prior_iq && pms == 0) || n_sv <= 0 || nd <= 0 || l0 < -100 ||
      l0 + nd - (type == 1 ? 1 : 0) > 100) {
    gnssb200_set_error(-7, "gnssb200_gpssdr_acquire: bad arguments (type 0, 1 or 2, Doppler range within +-100 kHz)", __FILE__, __LINE__);
    return -7;
  }
  for (int i = 0; i < n_sv; i++)
    if (sv_list[i] < 0 || sv_list[i] >= n_codes) {
      gnssb200_set_error(-7, "gnssb200_gpssdr_acquire: sv outside the code table", __FILE__, __LINE__);
      return -7;
    }
  CUDA_TRY(cudaSetDevice(h->device));
  std::vector<uint32_t> twf, twi, wipe;
  std::vector<int2> dft;
  make_twiddles(twf, twi);
  make_wipeoff(fif, wipe);
  make_dft(dft);
  const int per_sv = type == 2 ? nd * 8 : nd * 4, n_out = per_sv * n_sv;
  // rows the kernels may read: 4*ms of this preparation, 70 for doAcqMedium, those a prior preparation filled
  const int n_rows = std::max(std::max(4 * ms, 4 * pms), type == 1 ? 70 : 0);
  uint32_t *d_iq = nullptr, *d_piq = nullptr, *d_wipe = nullptr, *d_twf = nullptr, *d_twi = nullptr, *d_rows = nullptr, *d_codes = nullptr;
  int *d_sv = nullptr;
  Best *d_out = nullptr;
  std::vector<Best> out(n_out);
  cudaError_t e = cudaSuccess;
  auto A = [&](void **p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  auto U = [&](void *d, const void *s, size_t bytes) {
    if (e == cudaSuccess) e = cudaMemcpy(d, s, bytes, cudaMemcpyHostToDevice);
  };
  A((void **)&d_iq, (size_t)ms * NS * 4);
  if (pms) A((void **)&d_piq, (size_t)pms * NS * 4);
  A((void **)&d_wipe, wipe.size() * 4);
  A((void **)&d_twf, 4096);
  A((void **)&d_twi, 4096);
  A((void **)&d_rows, (size_t)n_rows * ROWLEN * 4);
  A((void **)&d_codes, (size_t)n_codes * NS * 4);
  A((void **)&d_sv, sizeof(int) * n_sv);
  A((void **)&d_out, sizeof(Best) * n_out);
  U(d_iq, iq, (size_t)ms * NS * 4);
  if (pms) U(d_piq, prior_iq, (size_t)pms * NS * 4);
  if (e == cudaSuccess && type == 1) e = cudaMemset(d_rows, 0, (size_t)n_rows * ROWLEN * 4);  // a fresh object: zeroed rows
  U(d_wipe, wipe.data(), wipe.size() * 4);
  U(d_twf, twf.data(), 4096);
  U(d_twi, twi.data(), 4096);
  U(d_codes, prn_codes, (size_t)n_codes * NS * 4);
  if (e == cudaSuccess) {
    int4 d4[100];
    for (int k = 0; k < 100; k++)
      d4[k] = make_int4((int16_t)(dft[k].x & 0xffff), (int16_t)((unsigned)dft[k].x >> 16), (int16_t)(dft[k].y & 0xffff), (int16_t)((unsigned)dft[k].y >> 16));
    e = cudaMemcpyToSymbol(c_dft, d4, sizeof d4);
  }
  U(d_sv, sv_list, sizeof(int) * n_sv);
  if (e == cudaSuccess) {
    cudaEventRecord(h->ev0, 0);
    if (pms) gsa_prep_kernel<<<4 * pms, NT>>>(d_piq, pms, d_wipe, d_twf, d_rows);
    gsa_prep_kernel<<<4 * ms, NT>>>(d_iq, ms, d_wipe, d_twf, d_rows);
    const size_t smem = ((size_t)20 * NS + SCR) * 4;
    if (type == 0)
      gsa_strong_kernel<<<n_out, NT>>>(d_rows, d_codes, d_sv, nd, l0, d_twi, d_out);
    else if (type == 1) {
      e = cudaFuncSetAttribute(gsa_weak_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess) gsa_weak_kernel<true><<<n_out, NT, smem>>>(d_rows, d_codes, d_sv, nd, l0, d_twi, d_out);
    } else {
      e = cudaFuncSetAttribute(gsa_weak_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess) gsa_weak_kernel<false><<<n_out, NT, smem>>>(d_rows, d_codes, d_sv, nd, l0, d_twi, d_out);
    }
    cudaEventRecord(h->ev1, 0);
    h->launches += pms ? 3 : 2;
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out.data(), d_out, sizeof(Best) * n_out, cudaMemcpyDeviceToHost);
  cudaFree(d_iq); cudaFree(d_piq); cudaFree(d_wipe); cudaFree(d_twf); cudaFree(d_twi); cudaFree(d_rows); cudaFree(d_codes); cudaFree(d_sv);
  cudaFree(d_out);
  if (e != cudaSuccess) {
    gnssb200_set_error((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    return (int)e;
  }
  // the reference's pick: combinations in loop order, a later one wins only if strictly larger
  for (int s = 0; s < n_sv; s++) {
    gnssb200_gpssdr_result r = {};
    int mag = 0;
    r.sv = sv_list[s];
    for (int c = 0; c < per_sv; c++) {
      const Best b = out[(size_t)s * per_sv + c];
      if (b.mag > mag) {
        mag = b.mag;
        if (type == 0) {
          const int l = l0 + c / 4, l2 = c % 4;
          r.code_phase = 2048 - b.idx;
          r.doppler = (int32_t)((l * 1000) + (float)l2 * 250);
        } else if (type == 1) {
          const int l = l0 + c / 4, l2 = c % 4;
          r.code_phase = b.idx % NS;
          r.doppler = (int32_t)((l * 1000) + (l2 * 250) + (b.idx / NS) * 25.0);
        } else {
          const int l = l0 + c / 8, l2 = (c / 2) % 4;
          r.code_phase = b.idx % NS;
          r.doppler = (int32_t)((l * 1000) + (l2 * 250) + (b.idx / NS) * 25.0);
        }
        r.magnitude = (uint32_t)mag;
      }
    }
    r.type = type;
    r.success = r.magnitude > 0u ? 1 : 0;  // THRESH_STRONG = THRESH_MEDIUM = THRESH_WEAK = 0 (RT/includes/config.h:72-74)
    results[s] = r;
  }
  return 0;
}

extern "C" int gnssb200_gpssdr_acquire(gnssb200_handle *h, const int16_t *iq, int type, double fif, const int16_t *prn_codes, int n_codes,
                                       const int32_t *sv_list, int n_sv, int doppmin, int doppmax, gnssb200_gpssdr_result *results) {
  if (type != 0 && type != 2) {
    gnssb200_set_error(-7, "gnssb200_gpssdr_acquire: type 0 (strong) or 2 (weak); medium is gnssb200_gpssdr_acquire_medium", __FILE__, __LINE__);
    return -7;
  }
  return gpssdr_run(h, iq, type, nullptr, 0, fif, prn_codes, n_codes, sv_list, n_sv, doppmin, doppmax, results);
}

extern "C" int gnssb200_gpssdr_acquire_medium(gnssb200_handle *h, const int16_t *iq, const int16_t *prior_iq, int prior_type, double fif,
                                              const int16_t *prn_codes, int n_codes, const int32_t *sv_list, int n_sv, int doppmin,
                                              int doppmax, gnssb200_gpssdr_result *results) {
  return gpssdr_run(h, iq, 1, prior_iq, prior_type, fif, prn_codes, n_codes, sv_list, n_sv, doppmin, doppmax, results);
}
