// navbits.cu -- first consumer of the tracking output: where the navigation message starts in each channel.
//
// What it replaces (SCI = trunk/GNSS_SOFTWARE_RECEIVERS/POSTPROCESSING_SCILAB_RECEIVERS):
//   GPS      [firstSubFrame, activeChnList] = findPreambles(...)   SCI/GPS/L1/findPreambles.sci:30-169
//            status = navPartyChk(ndat)                            SCI/GPS/L1/include/navPartyChk.sci:57-99
//   GLONASS  [firstString, activeChnList] = findTimeMarks(...)     SCI/GLONASS/L1/findTimeMarks.sci:25-66
// (the C receiver's own pream() is commented out: OSG/isr/osgpsisr.c:787-899).
//
// Both reference functions convolve sign(I_P) with a +-1 pattern held for 20 (GPS preamble, 8 bits) or 10
// (GLONASS time mark, 30 bits) milliseconds per bit and look for |result| above a threshold.  The pattern is
// piecewise constant, so with the prefix sums S of sign(I_P) every output is a signed sum of 8 (30)
// differences of S -- exact integer arithmetic (the reference's convol() goes through an FFT; its values are
// these integers up to rounding noise).  One CTA per channel reads the prompt values straight from the
// device buffers the tracking kernels wrote (doubles of gnssb200_softtrack, int32 dump records of
// gnssb200_track_run), builds S in a scratch row, scans all lags, and -- GPS -- verifies a candidate the way
// the reference does: another candidate exactly 6000 ms later and the parity of the TLM and HOW words
// (bit values summed over 20 ms each).
#include "common.cuh"

namespace {

struct NavArgs {
  const uint8_t *ip;        // element (ch, ms) at ip + ch*ch_stride + ms*ms_stride (bytes)
  long long ch_stride, ms_stride;
  int dtype;                // GNSSB200_NAV_F64 / GNSSB200_NAV_I32
  int n_ch, n_ms;
  int offset;               // searchStartOffset
  int nbits, rep;           // pattern: nbits values held rep ms each
  int thresh;               // |corr| > thresh
  int verify_gps;           // 1: 6000-ms repeat + parity of two words
  int tmpl[32];             // pattern in time order
  const int32_t *active;    // [n_ch] nonzero: channel was tracking
  int32_t *S;               // scratch [n_ch][n_ms + 1]
  int32_t *first;           // out [n_ch]: 1-based ms index, 0 = none
};

__device__ __forceinline__ double nav_value(const NavArgs &a, int ch, long long ms) {
  const uint8_t *p = a.ip + (long long)ch * a.ch_stride + ms * a.ms_stride;
  return a.dtype == GNSSB200_NAV_F64 ? *reinterpret_cast<const double *>(p) : (double)*reinterpret_cast<const int32_t *>(p);
}
__device__ __forceinline__ int nav_sign(double v) { return v > 0.0 ? 1 : (v < 0.0 ? -1 : 0); }  // Scilab sign(): 0 -> 0

// navPartyChk (SCI/GPS/L1/include/navPartyChk.sci:57-99) on values in {-1, 0, +1}; d[0..31] = ndat(1..32)
__device__ int nav_parity_check(const int *din) {
  int d[33];
  for (int i = 0; i < 32; i++) d[i + 1] = din[i];  // 1-based like the reference
  if (d[2] != 1)
    for (int i = 3; i <= 26; i++) d[i] = -d[i];
  int par[6];
  par[0] = d[1] * d[3] * d[4] * d[5] * d[7] * d[8] * d[12] * d[13] * d[14] * d[15] * d[16] * d[19] * d[20] * d[22] * d[25];
  par[1] = d[2] * d[4] * d[5] * d[6] * d[8] * d[9] * d[13] * d[14] * d[15] * d[16] * d[17] * d[20] * d[21] * d[23] * d[26];
  par[2] = d[1] * d[3] * d[5] * d[6] * d[7] * d[9] * d[10] * d[14] * d[15] * d[16] * d[17] * d[18] * d[21] * d[22] * d[24];
  par[3] = d[2] * d[4] * d[6] * d[7] * d[8] * d[10] * d[11] * d[15] * d[16] * d[17] * d[18] * d[19] * d[22] * d[23] * d[25];
  par[4] = d[2] * d[3] * d[5] * d[7] * d[8] * d[9] * d[11] * d[12] * d[16] * d[17] * d[18] * d[19] * d[20] * d[23] * d[24] * d[26];
  par[5] = d[1] * d[5] * d[7] * d[8] * d[10] * d[11] * d[12] * d[13] * d[15] * d[17] * d[21] * d[24] * d[25] * d[26];
  int same = 0;
  for (int i = 0; i < 6; i++) same += par[i] == d[27 + i] ? 1 : 0;
  return same == 6 ? -d[2] : 0;
}

__device__ __forceinline__ int nav_corr(const NavArgs &a, const int32_t *S, int L, int k) {
  int c = 0;
  int lo = min(k, L);
  int s_lo = S[lo];
  for (int q = 0; q < a.nbits; q++) {
    const int hi = min(k + a.rep * (q + 1), L);
    const int s_hi = S[hi];
    c += a.tmpl[q] * (s_hi - s_lo);
    s_lo = s_hi;
  }
  return c;
}

__global__ void __launch_bounds__(256) navbits_kernel(const NavArgs a) {
  const int ch = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  __shared__ int part[256];
  __shared__ int s_first;
  if (tid == 0) s_first = 0x7fffffff;
  const int L = a.n_ms - a.offset;  // values searched: I_P(ch, 1+offset : $)
  int32_t *S = a.S + (size_t)ch * (a.n_ms + 1);
  const bool on = a.active[ch] != 0 && L >= 1;
  if (on) {
    // prefix sums of sign(I_P): S[i] = sum of the first i values; each thread owns a contiguous segment
    const int seg = (L + nt - 1) / nt;
    const int i0 = min(tid * seg, L), i1 = min(i0 + seg, L);
    int acc = 0;
    for (int i = i0; i < i1; i++) acc += nav_sign(nav_value(a, ch, a.offset + i));
    part[tid] = acc;
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int t = 0; t < nt; t++) {
        const int v = part[t];
        part[t] = run;
        run += v;
      }
      S[0] = 0;
    }
    __syncthreads();
    acc = part[tid];
    for (int i = i0; i < i1; i++) {
      acc += nav_sign(nav_value(a, ch, a.offset + i));
      S[i + 1] = acc;
    }
  }
  __syncthreads();
  if (on) {
    for (int k = tid; k < L; k += nt) {
      const int c = nav_corr(a, S, L, k);
      if ((c < 0 ? -c : c) <= a.thresh) continue;
      const int idx1 = k + 1 + a.offset;  // the reference's 1-based index (find(...) + searchStartOffset)
      if (!a.verify_gps) {
        atomicMin(&s_first, idx1);
        continue;
      }
      // findPreambles.sci:104-141: a candidate exactly one subframe later, then the parity of TLM and HOW
      if (k + 6000 >= L) continue;
      const int c2 = nav_corr(a, S, L, k + 6000);
      if ((c2 < 0 ? -c2 : c2) <= a.thresh) continue;
      if (idx1 - 41 < 0 || idx1 + 1198 >= a.n_ms) continue;  // cannot happen: offset >= 40 and the later candidate is inside the record
      int bits[62];
      for (int g = 0; g < 62; g++) {
        double sum = 0.0;  // sum(bits, 'r') over the 20 values of one bit
        for (int i = 0; i < 20; i++) sum += nav_value(a, ch, (long long)idx1 - 41 + 20 * g + i);
        bits[g] = nav_sign(sum);
      }
      if (nav_parity_check(bits) != 0 && nav_parity_check(bits + 30) != 0) atomicMin(&s_first, idx1);
    }
  }
  __syncthreads();
  if (tid == 0) a.first[ch] = (on && s_first != 0x7fffffff) ? s_first : 0;
}

int nav_run(gnssb200_handle *h, NavArgs &a, int32_t *first_out, int32_t *active_out, const int32_t *active_in, void *cuda_stream) {
  if (!h || a.n_ch <= 0 || a.n_ms <= 0 || !a.ip || !first_out) {
    gnssb200_set_error(-5, "navbits: bad arguments", __FILE__, __LINE__);
    return -5;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int32_t *d_active = nullptr, *d_first = nullptr, *d_S = nullptr;
  CUDA_TRY(cudaMalloc(&d_active, sizeof(int32_t) * a.n_ch));
  CUDA_TRY(cudaMalloc(&d_first, sizeof(int32_t) * a.n_ch));
  CUDA_TRY(cudaMalloc(&d_S, sizeof(int32_t) * (size_t)a.n_ch * (a.n_ms + 1)));
  int32_t *tmp = new int32_t[a.n_ch];
  for (int i = 0; i < a.n_ch; i++) tmp[i] = active_in ? active_in[i] : 1;
  cudaError_t e = cudaMemcpyAsync(d_active, tmp, sizeof(int32_t) * a.n_ch, cudaMemcpyHostToDevice, st);
  a.active = d_active;
  a.first = d_first;
  a.S = d_S;
  if (e == cudaSuccess) {
    cudaEventRecord(h->ev0, st);
    navbits_kernel<<<a.n_ch, 256, 0, st>>>(a);
    cudaEventRecord(h->ev1, st);
    e = cudaGetLastError();
    h->launches += 1;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(first_out, d_first, sizeof(int32_t) * a.n_ch, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  delete[] tmp;
  cudaFree(d_active);
  cudaFree(d_first);
  cudaFree(d_S);
  if (e != cudaSuccess) {
    gnssb200_set_error((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    return (int)e;
  }
  if (active_out)
    for (int i = 0; i < a.n_ch; i++) active_out[i] = first_out[i] != 0 ? 1 : 0;
  return 0;
}

}  // namespace

extern "C" int gnssb200_find_preambles(gnssb200_handle *h, const void *d_ip, int dtype, int64_t ch_stride_bytes, int64_t ms_stride_bytes,
                                       int n_ch, int n_ms, const int32_t *active_in, int32_t *first_subframe, int32_t *active_out,
                                       void *cuda_stream) {
  NavArgs a = {};
  a.ip = (const uint8_t *)d_ip;
  a.ch_stride = ch_stride_bytes;
  a.ms_stride = ms_stride_bytes;
  a.dtype = dtype;
  a.n_ch = n_ch;
  a.n_ms = n_ms;
  a.offset = 5000;  // searchStartOffset, findPreambles.sci:34
  a.nbits = 8;
  a.rep = 20;
  a.thresh = 153;   // :91
  a.verify_gps = 1;
  // convol(preamble_ms, bits) with preamble_bits = [1 1 -1 1 -1 -1 -1 1] (:41) correlates with the reversed pattern
  static const int pb[8] = {1, 1, -1, 1, -1, -1, -1, 1};
  for (int q = 0; q < 8; q++) a.tmpl[q] = pb[7 - q];
  return nav_run(h, a, first_subframe, active_out, active_in, cuda_stream);
}

extern "C" int gnssb200_find_time_marks(gnssb200_handle *h, const void *d_ip, int dtype, int64_t ch_stride_bytes, int64_t ms_stride_bytes,
                                        int n_ch, int n_ms, const int32_t *active_in, int32_t *first_string, int32_t *active_out,
                                        void *cuda_stream) {
  NavArgs a = {};
  a.ip = (const uint8_t *)d_ip;
  a.ch_stride = ch_stride_bytes;
  a.ms_stride = ms_stride_bytes;
  a.dtype = dtype;
  a.n_ch = n_ch;
  a.n_ms = n_ms;
  a.offset = 0;     // searchStartOffset, findTimeMarks.sci:27
  a.nbits = 30;
  a.rep = 10;
  a.thresh = 290;   // :51
  a.verify_gps = 0;
  // tm_long = kron(-tm_bits, ones(1,10)) (:44-45), convolved: time-ordered pattern = reversed -tm_bits
  static const int tm[30] = {-1, 1, 1, -1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, -1, 1, 1, 1, -1, 1, 1, -1, -1, -1, 1, 1, 1, 1, 1};
  for (int q = 0; q < 30; q++) a.tmpl[q] = -tm[29 - q];
  return nav_run(h, a, first_string, active_out, active_in, cuda_stream);
}
