// isr_device.cuh -- the per-dump integer channel logic of the reference, executed by one lane per
// (stream, channel) right after the correlator reduction so the loop closes on the device.
//
// Follows OSG/isr/osgpsisr.c (OSG = trunk/GNSS_SOFTWARE_RECEIVERS/POSTPROCESSING_RECEIVERS/
// osgnss_next_step/src): gpsisr :360-408, ch_acq :424-459, ch_confirm :475-518, ch_pull_in :535-673,
// ch_track :692-768, rss :77-91, sqrt_newton :148-178, fix_atan2 :199-231, and the register
// accessors OSG/gp2021/gp2021.c:11-28,74-130.  C `long` is int64_t (LP64 semantics).
// Channels only touch their own registers, so each channel's lane owns a private register view.
#pragma once
#include "common.cuh"
#ifdef TRACK_PROFILE_ISR  // cycle stamps of the ISR lane's sections (build with EXTRA_NVCC_FLAGS="-DTRACK_PROFILE -DTRACK_PROFILE_ISR")
__device__ long long g_isr_t[8];
#define ISR_STAMP(i) { long long _c = clock64(); g_isr_t[i] += _c - _t0; _t0 = _c; }
#define ISR_T0 long long _t0 = clock64();
#else
#define ISR_STAMP(i)
#define ISR_T0
#endif


// The slice of REG_write / REG_read that belongs to one channel.
struct ChRegs {
  int w_prn;        // REG_write[ch*8+0]
  int w_carr_hi;    // +3
  int w_carr_lo;    // +4
  int w_code_hi;    // +5
  int w_code_lo;    // +6
  int w_epoch;      // +7
  int w_slew;       // REG_write[ch*8+0x84]
  int r_meas[8];    // REG_read[ch*8+0..7]   (1..6 TIC latches, 7 epoch)
  int r_acc[6];     // REG_read[ch*8+0x84+0..5]  IL QL IP QP IE QE
};

__device__ __forceinline__ int reg16(int v) { return (int)(uint16_t)v; }  // outpwd(): unsigned short

// Receiver constants plus what the device ISR derives from them once per launch (host side, track_launch).
struct DevCfg : gnssb200_cfg {
  int mult_i;  // clock_mult when it is an integer in (-1024, 1024), else 0
  int fast32;  // 1: the straight-line 32-bit form of the loop filters / NCO words applies (see dev_isr_loops)
};

__device__ __forceinline__ void dev_put_nco(int &hi, int &lo, long long freq, int bits, const DevCfg &c) {
  // gp2021.c:90-91 / 109-110: long w = freq << (32-N); w = w * (double)5 -> long; the two 16-bit halves of
  // the low word go to the registers.  For an integral multiplier and abs(w) < 2^40 the product is exact in
  // double, so the integer product is the same value -- and only its low 32 bits are kept, so for
  // frequencies inside int32 the whole thing is one 32-bit shift and multiply.
  const int sh = 32 - bits;
  if (c.mult_i != 0 && sh >= 0 && sh <= 8 && freq == (long long)(int)freq) {
    const unsigned w = ((unsigned)(int)freq << sh) * (unsigned)c.mult_i;
    hi = (int)(w >> 16);
    lo = (int)(w & 0xffffu);
    return;
  }
  long long w = freq << sh;
  const double mult = c.clock_mult;
  const long long im = (long long)mult;
  if ((double)im == mult && w > -(1ll << 40) && w < (1ll << 40) && im > -1024 && im < 1024)
    w = w * im;
  else
    w = (long long)((double)w * mult);
  hi = reg16((int)(w >> 16));
  lo = reg16((int)(w & 0xffff));
}
__device__ __forceinline__ void dev_ch_carrier(ChRegs &r, const DevCfg &c, long long f) {
  dev_put_nco(r.w_carr_hi, r.w_carr_lo, f, c.carrier_nco_bits, c);
}
__device__ __forceinline__ void dev_ch_code(ChRegs &r, const DevCfg &c, long long f) {
  dev_put_nco(r.w_code_hi, r.w_code_lo, f, c.code_nco_bits, c);
}

__device__ __forceinline__ long long dev_abs_trunc(long long v) {  // int abs() applied to a long
  int t = (int)v;
  return (long long)(t < 0 ? -t : t);
}
__device__ __forceinline__ int dev_sgn(long long v) { return v > 0 ? 1 : (v == 0 ? 0 : -1); }

__device__ __forceinline__ long long dev_mag(long long a, long long b) {  // rss()
  long long c = dev_abs_trunc(a), d = dev_abs_trunc(b);
  if (c == 0 && d == 0) return 0;
  return (c > d) ? (d >> 1) + c : (c >> 1) + d;
}

// sqrt_newton() (osgpsisr.c:148-178).  For every argument 0 < L < 2^31 -- the reference only passes
// sums of two squared shorts -- the Newton iteration of the reference returns
//        max { x : x*(x-1) <= L }
// (verified exhaustively over all 2^31 - 1 arguments against the reference's own loop, see
// tests/test_isr_math.py for the sampled regression).  So the value is taken from an approximate
// float square root and corrected with exact integer tests, without branches (the lone ISR lane pays
// the full latency of every taken branch); larger arguments run the literal iteration.
__device__ __noinline__ unsigned dev_isqrt_newton(long long L) {
  long long t, div;
  unsigned r = (unsigned)L;
  if (L & 0xFFFF0000LL)
    div = (L & 0xFF000000LL) ? 0x3FFF : 0x3FF;
  else
    div = (L & 0x0FF00LL) ? 0x3F : ((L > 4) ? 0x7 : L);
  for (;;) {
    t = L / div + div;
    div = t >> 1;
    div += t & 1;
    if ((long long)r > div)
      r = (unsigned)div;
    else {
      if (1 / r == r - 1 && 1 % r == 0) r--;
      return r;
    }
  }
}
// Branch-free cores: each returns a value that is exact whenever it leaves `bad` untouched, and sets
// `bad` when its float estimate was not within +-1 (cannot happen for the documented ranges, but the
// callers then recompute with the literal slow forms, so exactness never rests on a float error bound).
// Keeping the checks out of line lets the five discriminator chains of one dump share a basic block.
__device__ __forceinline__ unsigned core_isqrt31(unsigned l, bool &bad) {  // 0 <= l < 2^31; 0 -> 0 (reference: L <= 0)
  float sf;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sf) : "f"((float)l));
  unsigned x = (unsigned)(sf + 0.5f);  // within +-1 of the answer; x <= 46342, so x*(x+1) < 2^32
  x = x == 0 ? 1u : x;
  x -= (x * (x - 1) > l) ? 1u : 0u;
  x += ((x + 1) * x <= l) ? 1u : 0u;
  bad |= (x * (x - 1) > l) | ((x + 1) * x <= l);
  return l == 0 ? 0u : x;
}
__device__ __forceinline__ unsigned dev_isqrt(long long L) {
  if (L <= 0) return 0;
  if (L < (1ll << 31)) {
    bool bad = false;
    unsigned x = core_isqrt31((unsigned)L, bad);
    if (bad) {
      const unsigned l = (unsigned)L;
      x = 1;
      while ((x + 1) * x <= l) x++;
    }
    return x;
  }
  return dev_isqrt_newton(L);
}

// fix_atan2() (osgpsisr.c:199-231), 1 rad = 16384.  The divisor is always the operand of larger
// magnitude, so abs(n) <= 2^14 and the cubic correction fits 32-bit arithmetic.
__device__ __forceinline__ int dev_atan2_n3(int n) {  // ((((n*n)>>14)*n)>>13)/9 with abs(n) <= 2^14
  return ((((n * n) >> 14) * n) >> 13) / 9;
}
// trunc((a << 14) / B) for 0 <= a <= B, 0 < B < 2^30 (so the quotient is <= 2^14): float estimate
// (error < 1) corrected once in each direction with the exact remainder.
__device__ __forceinline__ unsigned core_udiv_q14(unsigned a, unsigned B, bool &bad) {
  unsigned q = (unsigned)__fdividef((float)a * 16384.0f, (float)B);
  // true remainder lies in (-B, 2B), inside int32: wrapping 32-bit arithmetic recovers it exactly
  int rem = (int)((a << 14) - q * B);
  const bool lo = rem < 0;
  q -= lo ? 1u : 0u;
  rem += lo ? (int)B : 0;
  const bool hi = rem >= (int)B;
  q += hi ? 1u : 0u;
  rem -= hi ? (int)B : 0;
  bad |= (rem < 0) | (rem >= (int)B) | (B >= (1u << 30));
  return q;
}

// trunc(num / den) for abs(num) < 2^30, 0 < den < 2^20 and abs(quotient) < 2^20 (C semantics: toward zero)
__device__ __forceinline__ int core_div_small(int num, int den, bool &bad) {
  const unsigned an = (unsigned)(num < 0 ? -num : num);
  unsigned q = (unsigned)__fdividef((float)an, (float)den);
  int rem = (int)(an - q * (unsigned)den);
  const bool lo = rem < 0;
  q -= lo ? 1u : 0u;
  rem += lo ? den : 0;
  const bool hi = rem >= den;
  q += hi ? 1u : 0u;
  rem -= hi ? den : 0;
  bad |= (rem < 0) | (rem >= den);
  return num < 0 ? -(int)q : (int)q;
}
__device__ __forceinline__ int dev_div_small(int num, int den) {
  bool bad = false;
  const int q = core_div_small(num, den, bad);
  return bad ? num / den : q;
}

// fix_atan2() for arguments that fit int32 (always the case for the dumps of the reference: they are
// shorts and products of shorts shifted right by 8): same case split, same truncations, 32-bit
// arithmetic, one division, selects instead of the four-way branch.
//   x > 0, x >= abs(y):   n = (y<<14)/x,  n - n3
//   x <= 0, -x >= abs(y): n = (y<<14)/x,  n - n3 (+pi if y > 0, else -pi)
//   y > 0, y > abs(x):    n = (x<<14)/y,  pi/2 - n + n3
//   y < 0, -y > abs(x):   n = (x<<14)/y,  -n + n3 - pi/2
__device__ __forceinline__ int atan2_from_quotient(int y, int x, bool horiz, unsigned q) {
  const int half_pi = 25736, pi = 51472;
  const int n = ((y < 0) != (x < 0)) ? -(int)q : (int)q;
  const int t = n - dev_atan2_n3(n);
  const int rh = x > 0 ? t : (y > 0 ? t + pi : t - pi);
  const int rv = y > 0 ? half_pi - t : -t - half_pi;
  const int res = horiz ? rh : rv;
  return (x | y) == 0 ? 0 : res;
}
__device__ __forceinline__ int core_atan2_i32(int y, int x, bool &bad) {
  const unsigned ay = (unsigned)(y < 0 ? -y : y), ax = (unsigned)(x < 0 ? -x : x);
  const bool horiz = ax >= ay;  // cases 1, 2 (and x == y == 0)
  const unsigned num = horiz ? ay : ax, den = horiz ? ax : ay;
  const unsigned q = core_udiv_q14(num, den == 0 ? 1u : den, bad);
  return atan2_from_quotient(y, x, horiz, q);
}
__device__ __noinline__ int dev_atan2_i32_slow(int y, int x) {  // same, the division done in 64 bits
  const unsigned ay = (unsigned)(y < 0 ? -y : y), ax = (unsigned)(x < 0 ? -x : x);
  const bool horiz = ax >= ay;
  const unsigned num = horiz ? ay : ax, den = horiz ? ax : ay;
  const unsigned q = (unsigned)(((unsigned long long)num << 14) / (den == 0 ? 1u : den));
  return atan2_from_quotient(y, x, horiz, q);
}
__device__ __forceinline__ int dev_atan2_i32(int y, int x) {
  bool bad = false;
  const int r = core_atan2_i32(y, x, bad);
  return bad ? dev_atan2_i32_slow(y, x) : r;
}

// indices into gnssb200_chan.accum[]
#define A_IP 0
#define A_QP 1
#define A_IL 2
#define A_QL 3
#define A_IE 4
#define A_QE 5

// reference words of the channel's system: the GPS ones of osgpsisr.c, or the GLONASS hooks of correlator.c:116-118
__device__ __forceinline__ long long dev_carrier_ref(const gnssb200_chan &k, const DevCfg &c) { return k.system ? c.glonass_carrier_ref : c.gps_carrier_ref; }
__device__ __forceinline__ long long dev_code_ref(const gnssb200_chan &k, const DevCfg &c) { return k.system ? c.glonass_code_ref : c.gps_code_ref; }

__device__ __forceinline__ void dev_isr_search(gnssb200_chan &k, ChRegs &r, const DevCfg &c) {
  if (abs(k.n_freq) <= k.search_max_f) {
    long long pm = dev_mag(k.accum[A_IP], k.accum[A_QP]);
    if (pm > c.acq_thresh) {
      k.state = 2;
      k.i_confirm = 0;
      k.n_thresh = 0;
      k.mean_early = k.mean_prompt = k.mean_late = 0;
    } else {
      r.w_slew = reg16(1);
      k.codes += 1;
    }
    if (k.codes == k.search_max_PRN_delay) {
      k.n_freq += k.del_freq;
      k.del_freq = -(k.del_freq + dev_sgn(k.del_freq));
      k.carrier_freq = dev_carrier_ref(k, c) + k.carrier_cold_corr + c.d_freq * k.n_freq;
      dev_ch_carrier(r, c, k.carrier_freq);
      k.codes = 0;
    }
  } else {
    k.n_freq = 0;
    k.del_freq = 1;
    k.carrier_freq = dev_carrier_ref(k, c) + k.carrier_cold_corr + c.d_freq * k.n_freq;
    dev_ch_carrier(r, c, k.carrier_freq);
    k.codes = 0;
  }
  k.CN0 = 0;
}

__device__ __forceinline__ void dev_isr_confirm(gnssb200_chan &k, ChRegs &r, const DevCfg &c) {
  long long pm = dev_mag(k.accum[A_IP], k.accum[A_QP]);
  long long lm = dev_mag(k.accum[A_IL], k.accum[A_QL]);
  long long em = dev_mag(k.accum[A_IE], k.accum[A_QE]);
  k.mean_early += em;
  k.mean_prompt += pm;
  k.mean_late += lm;
  if (pm > c.acq_thresh) k.n_thresh++;
  if (k.i_confirm == 3) {          // CONFIRM_M
    if (k.n_thresh >= 2) {         // N_OF_M_THRESH
      k.state = 3;
      k.CN0 = 0;
      k.ch_time = 0;
      k.ms_set = 0;
      k.oldCarrNco = k.oldCodeNco = k.oldCarrError = k.oldCodeError = 0;
      k.codeFreqBasis = dev_code_ref(k, c);
      k.carrFreqBasis = k.carrier_freq;
      k.sign_pos = k.prev_sign_pos = 0;
    } else
      k.state = 1;
  }
  k.i_confirm++;
}

__device__ __forceinline__ void dev_isr_loops(gnssb200_chan &k, ChRegs &r, const DevCfg &c) {
  const int ip = k.accum[A_IP], qp = k.accum[A_QP], pip = k.prev_accum[A_IP], pqp = k.prev_accum[A_QP];
  const int ie = k.accum[A_IE], qe = k.accum[A_QE], il = k.accum[A_IL], ql = k.accum[A_QL];
  ISR_T0
  // The four discriminator primitives are evaluated up front and unconditionally (they are branch-free
  // and total), so their dependency chains interleave on the single ISR lane; the reference's
  // conditions select the results below.
  const int cross8 = (ip * pqp - pip * qp) >> 8;  // int products, as in the reference
  const int dt = ip * pip + qp * pqp;
  const int dot8 = (int)((dt < 0 ? 0u - (unsigned)dt : (unsigned)dt) >> 8);  // abs() of the widened value, as the reference
  // operands: abs(cross), dot < 2^23 and abs(qp), abs(ip) <= 2^15 (shorts) -> the int32 form is exact
  bool bad = false;
  int at_f = core_atan2_i32(cross8, dot8, bad);
  const int py = qp * dev_sgn(ip), px = ip < 0 ? -ip : ip;
  int at_p = core_atan2_i32(py, px, bad);
  const int Le = ie * ie + qe * qe, Ll = il * il + ql * ql;  // <= 2^31: a wrapped (negative) sum gives 0 like the reference
  unsigned se = core_isqrt31(Le < 0 ? 0u : (unsigned)Le, bad), sl = core_isqrt31(Ll < 0 ? 0u : (unsigned)Ll, bad);
  // (8192*(se-sl))/(se+sl): se, sl <= 46341, so everything fits int32 (C division truncates toward zero)
  int den_c = (int)se + (int)sl;
  int code_q = core_div_small(8192 * ((int)se - (int)sl), den_c > 0 ? den_c : 1, bad);  // abs(num) <= 2^29.5, den <= 92684, abs(q) <= 8192
  const bool carr_ok = ip != 0 && qp != 0 && pip != 0 && pqp != 0;
  const bool code_ok = ie != 0 && qe != 0 && il != 0 && ql != 0;
  const long long oce = k.oldCarrError, ode = k.oldCodeError;
  // Straight-line 32-bit form of both loop filters and both NCO words.  Valid when the carried-over errors
  // are small (they came out of fix_atan2 / the DLL discriminator: abs < 2^17), the filter coefficients keep
  // the numerators inside int32 (c.fast32, checked on the host: abs(e) < 2^17, abs(f) < 2^16, abs(codeError)
  // < 2^17), the clock multiplier is a small integer and the new frequencies fit int32 (checked below).
  if (c.fast32 && !bad && (unsigned long long)(oce + (1ll << 17)) < (1ull << 18) && (unsigned long long)(ode + (1ll << 17)) < (1ull << 18)) {
    const int e = carr_ok ? at_p / 2 : (int)oce;
    const int f = carr_ok ? at_f : 0;
    const int numc = c.pll_i1 * e - c.pll_i2 * (int)oce - c.pll_i3 * f;
    const long long carrNco = k.oldCarrNco + (long long)(numc / 51472);
    const long long carrFreq = k.carrFreqBasis + carrNco;
    const int d = code_ok ? code_q : (int)ode;
    const int numk = (c.dll_i1 + 1) * d - c.dll_i2 * (int)ode;
    const long long codeNco = k.oldCodeNco + (long long)(numk / 8192);
    const long long codeFreq = k.codeFreqBasis - codeNco;
    if (carrFreq == (long long)(int)carrFreq && codeFreq == (long long)(int)codeFreq) {
      const unsigned wc = ((unsigned)(int)carrFreq << (32 - c.carrier_nco_bits)) * (unsigned)c.mult_i;
      const unsigned wk = ((unsigned)(int)codeFreq << (32 - c.code_nco_bits)) * (unsigned)c.mult_i;
      r.w_carr_hi = (int)(wc >> 16);
      r.w_carr_lo = (int)(wc & 0xffffu);
      r.w_code_hi = (int)(wk >> 16);
      r.w_code_lo = (int)(wk & 0xffffu);
      if (carr_ok) {
        k.cross = (long long)cross8;
        k.dot = (long long)dot8;
      }
      k.freqError = (long long)f;
      k.carrError = (long long)e;
      k.carrNco = carrNco;
      k.oldCarrNco = carrNco;
      k.oldCarrError = (long long)e;
      k.carrFreq = carrFreq;
      k.codeError = (long long)d;
      k.codeNco = codeNco;
      k.oldCodeNco = codeNco;
      k.oldCodeError = (long long)d;
      k.codeFreq = codeFreq;
      return;
    }
  }
  if (bad) {  // an estimate was off by more than one (not expected): literal forms
    at_f = dev_atan2_i32_slow(cross8, dot8);
    at_p = dev_atan2_i32_slow(py, px);
    se = dev_isqrt((long long)Le);
    sl = dev_isqrt((long long)Ll);
    den_c = (int)se + (int)sl;
    code_q = (8192 * ((int)se - (int)sl)) / (den_c > 0 ? den_c : 1);
  }
  ISR_STAMP(0)
  const long long ce = carr_ok ? (long long)(at_p / 2) : oce;
  const long long fe = carr_ok ? (long long)at_f : 0ll;
  if (carr_ok) {
    k.cross = (long long)cross8;
    k.dot = (long long)dot8;
  }
  k.freqError = fe;
  k.carrError = ce;
  {
    // 32 x 32 -> 64-bit products when the carried-over error is an int32 (it always is: it came from fix_atan2)
    const long long num = (oce == (long long)(int)oce)
                              ? (long long)c.pll_i1 * (int)ce - (long long)c.pll_i2 * (int)oce - (long long)c.pll_i3 * (int)fe
                              : c.pll_i1 * ce - c.pll_i2 * oce - c.pll_i3 * fe;
    const long long q = (num == (long long)(int)num) ? (long long)((int)num / 51472) : num / 51472;
    k.carrNco = k.oldCarrNco + q;
  }
  k.oldCarrNco = k.carrNco;
  k.oldCarrError = ce;
  k.carrFreq = k.carrFreqBasis + k.carrNco;
  ISR_STAMP(1)
  dev_ch_carrier(r, c, k.carrFreq);
  ISR_STAMP(2)

  const long long de = code_ok ? (long long)code_q : ode;
  k.codeError = de;
  {
    const long long num = (ode == (long long)(int)ode) ? (long long)(c.dll_i1 + 1) * (int)de - (long long)c.dll_i2 * (int)ode
                                                       : (c.dll_i1 + 1) * de - c.dll_i2 * ode;
    const long long q = (num == (long long)(int)num) ? (long long)((int)num / 8192) : num / 8192;
    k.codeNco = k.oldCodeNco + q;
  }
  k.oldCodeNco = k.codeNco;
  k.oldCodeError = de;
  k.codeFreq = k.codeFreqBasis - k.codeNco;
  ISR_STAMP(3)
  dev_ch_code(r, c, k.codeFreq);
  ISR_STAMP(4)
}

// Pull-in and tracking are split in two: the part that decides the channel's NCO words (what the next
// block's correlator parameters depend on) and the bookkeeping that only touches the channel struct, so
// the kernel can publish the next block's parameters before the bookkeeping runs.
__device__ __forceinline__ void dev_isr_pull_in_words(gnssb200_chan &k, ChRegs &r, const DevCfg &c) {
  dev_isr_loops(k, r, c);
  if (k.ch_time + 1 == 3000) {  // the time-out below will reload the reference words
    dev_ch_carrier(r, c, dev_carrier_ref(k, c));
    dev_ch_code(r, c, dev_code_ref(k, c));
  }
}
__device__ __forceinline__ void dev_isr_pull_in_rest(gnssb200_chan &k, ChRegs &r, const DevCfg &c) {
  const int ip = k.accum[A_IP], pip = k.prev_accum[A_IP];
  ISR_T0
  if (dev_sgn(ip) == -dev_sgn(pip)) {
    k.prev_sign_pos = k.sign_pos;
    k.sign_pos = (int)k.ch_time;
    if (k.sign_pos - k.prev_sign_pos > 19)
      k.sign_count++;
    else
      k.sign_count = 0;
  }
  k.ms_count++;
  if ((dev_sgn(ip) == -1 && (k.ms_sign & 0xfffffULL) == 0x00000ULL) ||
      (dev_sgn(ip) == 1 && (k.ms_sign & 0xfffffULL) == 0xfffffULL)) {
    if (dev_sgn(ip) == -dev_sgn(pip)) {
      k.ms_count = 0;
      r.w_epoch = reg16(0x1);
      k.ms_set = 1;
    }
  }
  k.ms_sign <<= 1;
  if (ip < 0) k.ms_sign |= 1;
  k.ms_count %= 20;
  k.ch_time++;
  if (k.sign_count > 30 && k.ms_set) k.state = 4;
  if (k.ch_time == 3000) {
    k.del_freq = 1;
    k.n_freq = 0;
    dev_ch_carrier(r, c, dev_carrier_ref(k, c));  // same words as dev_isr_pull_in_words wrote
    dev_ch_code(r, c, dev_code_ref(k, c));
    k.codes = 0;
    k.ch_time = 0;
    k.state = 1;
  }
  ISR_STAMP(5)
}

__device__ __forceinline__ void dev_isr_track_rest(gnssb200_chan &k) {
  k.ms_count = (k.ms_count + 1) % 20;
  if (k.ms_count == 19) k.bit = k.accum[A_IP] > 0 ? 1 : 0;
}

// One channel's share of gpsisr() for a block in which it dumped, first part: the new dump values and
// everything that can change the channel's write registers (NCO words, slew).  state_in receives the
// state the dump was handled in.  Returns 1 on CHANNEL_OFF (the reference exit(0)s there).
__device__ __forceinline__ int dev_gpsisr_words(gnssb200_chan &k, ChRegs &r, const DevCfg &c, int &state_in) {
#pragma unroll
  for (int a = 0; a < 6; a++) k.prev_accum[a] = k.accum[a];
  k.accum[A_IE] = (int16_t)r.r_acc[4];  // from_gps(): short
  k.accum[A_QE] = (int16_t)r.r_acc[5];
  k.accum[A_IP] = (int16_t)r.r_acc[2];
  k.accum[A_QP] = (int16_t)r.r_acc[3];
  k.accum[A_IL] = (int16_t)r.r_acc[0];
  k.accum[A_QL] = (int16_t)r.r_acc[1];
  state_in = k.state;
  switch (k.state) {
    case 0: return 1;
    case 1: dev_isr_search(k, r, c); break;
    case 3: dev_isr_pull_in_words(k, r, c); break;
    case 4: dev_isr_loops(k, r, c); break;
    default: break;
  }
  return 0;
}
// second part: bookkeeping on the channel struct (and the epoch-load request of the bit sync)
__device__ __forceinline__ void dev_gpsisr_rest(gnssb200_chan &k, ChRegs &r, const DevCfg &c, int state_in) {
  switch (state_in) {
    case 2: dev_isr_confirm(k, r, c); break;
    case 3: dev_isr_pull_in_rest(k, r, c); break;
    case 4: dev_isr_track_rest(k); break;
    default: break;
  }
}

// Both parts back to back (the order of the reference).
__device__ __forceinline__ int dev_gpsisr_channel(gnssb200_chan &k, ChRegs &r, const DevCfg &c) {
  int st;
  if (dev_gpsisr_words(k, r, c, st)) return 1;
  dev_gpsisr_rest(k, r, c, st);
  return 0;
}
