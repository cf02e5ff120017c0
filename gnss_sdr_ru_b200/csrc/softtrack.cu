// softtrack.cu -- the Scilab receivers' floating-point tracking:
// [trackResults, channel] = tracking(fid, channel, settings),
// SCI/GLONASS/L1/tracking.sci:226-400 and SCI/GPS/L1/tracking.sci (SCI = trunk/GNSS_SOFTWARE_RECEIVERS/
// POSTPROCESSING_SCILAB_RECEIVERS): one code period per iteration, block size ceil((L-rem)/step),
// float carrier / code NCOs, early/prompt/late sums, FLL-assisted PLL and DLL in double precision.
//
// One CTA per channel (channels are independent and each is sequential in time).  Every sample of a
// block is independent given the block's NCO state, so threads stride over the block (coalesced int8
// loads), evaluate the carrier with a double-precision sincos of the reference's own argument
// expression, index the code with ceil() exactly like the reference, and six double sums are reduced
// by shuffles; lane 0 runs the discriminators / loop filters and publishes the next block's state.
// FP64 throughout: the loop is a feedback system in doubles in the reference, FP32 would not track it.
#include <math.h>

#include <vector>

#include "common.cuh"

struct FtrkChanDev {
  int sv;               // PRN (GPS) or frequency channel (GLONASS)
  int code_row;         // row in the chip table (0: GLONASS ST code, 1..32: C/A)
  double acquired_freq;
  long long start;      // first complex sample of the channel (skip + codePhase - 1)
};

struct FtrkArgs {
  const int8_t *iq;
  long long n_samples;
  const int8_t *chips;  // [33][1024]
  const FtrkChanDev *chan;
  double *out;          // [n_ch][ms][13]
  int32_t *ms_done;     // [n_ch]
  int ms;
  int glonass;
  int code_len;
  double fs, IF, IF_step, zero_channel, code_freq_basis, spc;
  double k1, k2, k3, tau1, tau2;
};

struct FtrkState {
  double codeFreq, remCodePhase, carrFreq, remCarrPhase;
  long long pos;
  int blksize;
  int stop;
};

__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(512) softtrack_kernel(const FtrkArgs a) {
  __shared__ double code[1023 + 2];
  __shared__ FtrkState st;
  __shared__ double red[16][6];
  const int ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const FtrkChanDev c = a.chan[ch];
  const int L = a.code_len;
  // caCode = [caCode($) caCode caCode(1)]  (tracking.sci:171-175)
  for (int i = tid; i < L + 2; i += blockDim.x) {
    const int k = (i == 0) ? L - 1 : (i == L + 1 ? 0 : i - 1);
    code[i] = (double)a.chips[c.code_row * 1024 + k];
  }
  // loop state that only lane 0 needs
  double oldCodeNco = 0.0, oldCodeError = 0.0, oldCarrNco = 0.0, oldCarrError = 0.0;
  double I1 = 0.001, Q1 = 0.001;
  const double carrFreqBasis = c.acquired_freq;
  if (tid == 0) {
    st.codeFreq = a.code_freq_basis;
    st.remCodePhase = 0.0;
    st.carrFreq = c.acquired_freq;
    st.remCarrPhase = 0.0;
    st.pos = c.start;
    st.stop = 0;
    const double step = st.codeFreq / a.fs;
    st.blksize = (int)ceil(((double)L - st.remCodePhase) / step);
    if (st.pos < 0 || st.pos + st.blksize > a.n_samples) st.stop = 1;
  }
  __syncthreads();
  double *out = a.out + (size_t)ch * a.ms * 13;
  int done = 0;
  for (int it = 0; it < a.ms; it++) {
    if (st.stop) break;
    const double codeFreq = st.codeFreq, rem = st.remCodePhase, carrFreq = st.carrFreq, remCarr = st.remCarrPhase;
    const long long pos = st.pos;
    const int blksize = st.blksize;
    const double step = codeFreq / a.fs;            // codePhaseStep
    const double w = (carrFreq * 2.0) * M_PI;       // (carrFreq * 2.0 * %pi)
    const double remE = rem - a.spc, remL = rem + a.spc;
    double s[6] = {0, 0, 0, 0, 0, 0};               // I_E Q_E I_P Q_P I_L Q_L
    const char2 *src = reinterpret_cast<const char2 *>(a.iq) + pos;
    // Carrier replica exp(i*trigarg), trigarg = ((carrFreq*2*pi) .* time) + remCarrPhase (tracking.sci:288-291).
    // A thread's samples are blockDim.x apart, i.e. a fixed phase step apart: one double-precision sincos of
    // the reference's own argument for its first sample, one for the step, then a complex rotation per sample
    // (4 FP64 multiply-adds instead of a ~100-instruction sincos; 31 rotations add ~1e-15 of relative error, the
    // reference's own argument carries ~1e-12 rad of rounding at 6000 rad).
    double sn = 0.0, cs = 1.0, sd, cd;
    sincos(__dmul_rn(w, __ddiv_rn((double)blockDim.x, a.fs)), &sd, &cd);
    if (tid < blksize) sincos(__dadd_rn(__dmul_rn(w, __ddiv_rn((double)tid, a.fs)), remCarr), &sn, &cs);
    for (int j = tid; j < blksize; j += blockDim.x) {
      const char2 v = __ldg(src + j);
      const double I = (double)v.x, Q = (double)v.y;
      // carrsig .* rawSignal, carrsig = exp(%i*trigarg); qBaseband = real, iBaseband = imag
      const double qb = __dsub_rn(__dmul_rn(cs, I), __dmul_rn(sn, Q));
      const double ib = __dadd_rn(__dmul_rn(cs, Q), __dmul_rn(sn, I));
      const double dj = (double)j;
      const double e = code[(int)ceil(__dadd_rn(remE, __dmul_rn(dj, step)))];   // caCode(ceil(tcode)+1), 1-based
      const double p = code[(int)ceil(__dadd_rn(rem, __dmul_rn(dj, step)))];
      const double l = code[(int)ceil(__dadd_rn(remL, __dmul_rn(dj, step)))];
      s[0] += e * ib;
      s[1] += e * qb;
      s[2] += p * ib;
      s[3] += p * qb;
      s[4] += l * ib;
      s[5] += l * qb;
      const double c2 = cs * cd - sn * sd, s2 = sn * cd + cs * sd;  // advance the replica by blockDim.x samples
      cs = c2;
      sn = s2;
    }
#pragma unroll
    for (int q = 0; q < 6; q++) {
      const double r = warp_sum_d(s[q]);
      if (lane == 0) red[warp][q] = r;
    }
    __syncthreads();
    if (tid == 0) {
      double t[6];
      for (int q = 0; q < 6; q++) {
        double acc = 0.0;
        for (int wv = 0; wv < nwarps; wv++) acc += red[wv][q];
        t[q] = acc;
      }
      const double I_E = t[0], Q_E = t[1], I_P = t[2], Q_P = t[3], I_L = t[4], Q_L = t[5];
      // tracking.sci:301-303, 309-312
      const double newRem = __dadd_rn(__dadd_rn(rem, __dmul_rn((double)(blksize - 1), step)), step) - (double)L;
      const double last = __dadd_rn(__dmul_rn(w, __ddiv_rn((double)blksize, a.fs)), remCarr);
      const double newRemCarr = __dsub_rn(last, __dmul_rn(trunc(last / (2 * M_PI)), 2 * M_PI));
      // FLL-assisted PLL (:327-347)
      const double I2 = I1, Q2 = Q1;
      I1 = I_P;
      Q1 = Q_P;
      const double cross = I1 * Q2 - I2 * Q1;
      const double dot = fabs(I1 * I2 + Q1 * Q2);
      const double freqError = atan2(cross, dot) / M_PI;
      const double carrError = atan(Q_P / I_P) / (2.0 * M_PI);
      const double carrNco = oldCarrNco + a.k1 * carrError - a.k2 * oldCarrError - a.k3 * freqError;
      oldCarrNco = carrNco;
      oldCarrError = carrError;
      const double newCarrFreq = carrFreqBasis + carrNco;
      // DLL (:352-371)
      const double sE = sqrt(I_E * I_E + Q_E * Q_E), sL = sqrt(I_L * I_L + Q_L * Q_L);
      const double codeError = (sE - sL) / (sE + sL);
      const double codeNco = oldCodeNco + (a.tau2 / a.tau1) * (codeError - oldCodeError) + codeError * (0.001 / a.tau1);
      oldCodeNco = codeNco;
      oldCodeError = codeError;
      double newCodeFreq;
      if (a.glonass)
        newCodeFreq = a.code_freq_basis - codeNco +
                      (newCarrFreq - (a.IF + a.IF_step * c.sv)) / ((a.zero_channel + c.sv * a.IF_step) / a.code_freq_basis);
      else
        newCodeFreq = a.code_freq_basis - codeNco + ((newCarrFreq - a.IF) / 1540);
      const long long newPos = pos + blksize;
      double *o = out + (size_t)it * 13;
      o[0] = I_E; o[1] = I_P; o[2] = I_L; o[3] = Q_E; o[4] = Q_P; o[5] = Q_L;
      o[6] = newCarrFreq; o[7] = newCodeFreq; o[8] = codeError; o[9] = codeNco; o[10] = carrError; o[11] = carrNco;
      o[12] = (double)newPos - newRem * (a.fs / 1000) / (double)L;   // absoluteSample (:380-384)
      st.codeFreq = newCodeFreq;
      st.remCodePhase = newRem;
      st.carrFreq = newCarrFreq;
      st.remCarrPhase = newRemCarr;
      st.pos = newPos;
      const double nstep = newCodeFreq / a.fs;
      st.blksize = (int)ceil(((double)L - newRem) / nstep);
      if (newPos + st.blksize > a.n_samples) st.stop = 1;
    }
    done = it + 1;
    __syncthreads();
  }
  if (tid == 0) a.ms_done[ch] = done;
}

extern "C" int gnssb200_softtrack(gnssb200_handle *h, const gnssb200_softtrack_cfg *cfg, const void *d_iq, int64_t n_samples,
                                  const gnssb200_softtrack_chan *chans, int n_ch, double *d_out, int32_t *d_ms_done,
                                  void *cuda_stream) {
  if (!h || !cfg || !d_iq || !chans || n_ch <= 0 || !d_out || !d_ms_done || cfg->ms_to_process <= 0 ||
      (cfg->code_length != 511 && cfg->code_length != 1023)) {  // the kernel's chip buffer holds one ST or C/A period
    gnssb200_set_error(-20, "gnssb200_softtrack: bad arguments", __FILE__, __LINE__);
    return -20;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int8_t *d_chips = nullptr;
  if (int rc = chip_table(h, &d_chips)) return rc;
  const bool glo = cfg->system == GNSSB200_SYS_GLONASS;
  std::vector<FtrkChanDev> hc(n_ch);
  for (int i = 0; i < n_ch; i++) {
    hc[i].sv = chans[i].sv;
    hc[i].code_row = glo ? 0 : chans[i].sv;
    if (!glo && (chans[i].sv < 1 || chans[i].sv > 32)) {
      gnssb200_set_error(-21, "gnssb200_softtrack: GPS PRN must be 1..32", __FILE__, __LINE__);
      return -21;
    }
    hc[i].acquired_freq = chans[i].acquired_freq;
    hc[i].start = cfg->skip_samples + (int64_t)(chans[i].code_phase - 1);
  }
  FtrkChanDev *d_ch = nullptr;
  CUDA_TRY(cudaMalloc(&d_ch, sizeof(FtrkChanDev) * n_ch));
  CUDA_TRY(cudaMemcpyAsync(d_ch, hc.data(), sizeof(FtrkChanDev) * n_ch, cudaMemcpyHostToDevice, st));
  FtrkArgs a;
  a.iq = (const int8_t *)d_iq;
  a.n_samples = n_samples;
  a.chips = d_chips;
  a.chan = d_ch;
  a.out = d_out;
  a.ms_done = d_ms_done;
  a.ms = cfg->ms_to_process;
  a.glonass = glo ? 1 : 0;
  a.code_len = cfg->code_length;
  a.fs = cfg->samp_freq;
  a.IF = cfg->IF;
  a.IF_step = cfg->IF_step;
  a.zero_channel = cfg->glonass_zero_channel;
  a.code_freq_basis = cfg->code_freq;
  a.spc = cfg->dll_correlator_spacing;
  {  // calcLoopCoef.sci:38-43 (k = 1.0), calcFLLPLLLoopCoef.sci:36-38 (T = 0.001)
    const double zeta = cfg->dll_damping_ratio, LBW = cfg->dll_noise_bandwidth;
    const double Wn = LBW * 8 * zeta / (4 * zeta * zeta + 1);
    a.tau1 = 1.0 / (Wn * Wn);
    a.tau2 = 2.0 * zeta / Wn;
    const double T = 0.001, p = cfg->pll_noise_bandwidth / 0.53;
    a.k1 = T * (p * p) + 1.414 * p;
    a.k2 = 1.414 * p;
    a.k3 = T * (cfg->fll_noise_bandwidth / 0.25);
  }
  CUDA_TRY(cudaEventRecord(h->ev0, st));
  softtrack_kernel<<<n_ch, 512, 0, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(h->ev1, st));
  h->launches++;
  CUDA_TRY(cudaStreamSynchronize(st));  // hc / d_ch lifetimes
  cudaFree(d_ch);
  return 0;
}
