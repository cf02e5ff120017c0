// track.cu -- GP2021-style E/P/L correlator bank + closed channel loop on the device.
//
// What it replaces: Sim_GP2021_int (OSG/correlator/correlator.c:148-316) followed by gpsisr
// (OSG/isr/osgpsisr.c:360-408) once per 512 us block, for S independent IF streams x 12 channels.
//
// Two kernels share the arithmetic below (closed-form NCO starts, grouped code taps, dump rules, device ISR):
//
//   track_ws_kernel    the hot path (8192-sample blocks or shorter, 16-byte aligned, int8 I,Q or packed 2+2 bit):
//                      warp-specialised -- correlator warps + one control lane per CTA, decoupled by mbarriers,
//                      samples staged by the TMA engine, ISR bookkeeping after the next block's parameters are
//                      published -- and scheduled through a (channel, time-slice) work queue.  Described where
//                      it is defined.
//   track_loop_kernel  the generic variant (any block length, unaligned or ragged records, I-only input; also
//                      selectable with GNSSB200_TRACK_WS=0 for A/B runs): one CTA per (stream, channel) for the
//                      whole run, barrier-synchronised: TMA double buffering, 32 samples per thread per pass,
//                      warp-shuffle + shared-atomic reduction, lane 0 of warp 0 as the ISR lane.
//
// Per block, both: every thread takes consecutive complex samples, starts its carrier/code NCOs from the closed
// form phase(i) = phase0 + i*incr and walks them with the exact integer arithmetic of the
// reference: 8-phase LO (phase>>29), complex mix, +-1 E/P/L taps at half-chip spacing.  I and Q products ride in
// one 32-bit register as two 16-bit lanes.  Samples go four at a time: with 4*kinc < 2^32 at most one code-NCO
// carry falls inside a group, so the group contributes old_bits*S_old + new_bits*S_new -- 6 IMADs and one table
// read per 4 samples.  At most one dump per block on this path; the single straddling chunk is re-evaluated
// sample-per-lane by its warp from the closed forms.
//
// Register values the fast path cannot express (a second dump inside one block, slews beyond the table window,
// PRN outside 1..32, ...) take the serial path: one lane walks the block with the literal per-sample loop.  It is
// still device code; there is no CPU fallback.
#include "isr_device.cuh"

#define MODE_STOP (-1)
#define MODE_IDLE 0
#define MODE_FAST 1
#define MODE_SERIAL 2

// Shared-memory copy of the channel's code-table row plus the start of the next one (the reference's
// spill-over reads).  Table indices of a closed-form block stay below hc0 + w1 (= the dump position, 2046 +
// slew) before the dump and below the half chips one block spans (~1050) after it, each plus the few entries
// a chunk reads ahead; larger slews take the serial path.
#define SMEM_TBL 2304

struct StepParams {
  uint32_t cph0, kph0, cinc, kinc;
  uint32_t hc0, w1, stale_idx, slew_dump;
  int mode;
  int tic_count;
  uint32_t stale_bits;   // table entry at stale_idx
  uint32_t cyc_pending;  // carrier wraps of quiet blocks not yet added to gnssb200_corr.carrier_cycle
  long long tic;         // value of the TIC down-counter after this block's tic_count was derived
};

struct TrackArgs {
  gnssb200_rx *rx;
  int32_t *chan_flags;
  const uint32_t *code_table;
  const uint8_t *d_if;
  size_t stride;
  int fmt, nsamp;
  long long nblocks;
  int run_isr;
  int first_stream;
  gnssb200_dump *dumps;
  int dump_cap;
  int32_t *dump_count;
  DevCfg cfg;
  // 1, 8, 128, 2048 as run-time values: multiplying by them keeps address arithmetic of the hot loop on the
  // FMA pipe (IMAD) instead of the ALU pipe (SHF/LOP3/IADD3), which is the busier one (ptxas would turn a
  // multiplication by a literal power of two back into a shift)
  uint32_t k1, k8, k128, k2048;
  struct SchedQueue *sched;  // work queue of this launch (track_ws_kernel)
};

// byte k of w, sign extended, in one PRMT: selector nibble k copies the byte, nibble k|8 replicates
// its sign bit (PTX prmt default mode).  Inline PTX because __byte_perm() documents only 3 selector bits.
__device__ __forceinline__ int sext8(uint32_t w, int k) {
  const uint32_t sel = 0x8880u | (uint32_t)(k * 0x1111);
  int r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"(sel));
  return r;
}

// 8-phase LO of the GP2021 / Namuru carrier NCO (correlator.c:203-204, NAM/rtl/carrier_nco.v:21-25),
// stored so that  I*lut.x + Q*lut.y = ival + 65536*qval  with
//   ival = i_lo*I + q_lo*Q,  qval = q_lo*I - i_lo*Q   (correlator.c:214-215)
__device__ __forceinline__ void fill_lo_lut(uint2 *lut) {
  const int i_lo[8] = {-1, 1, 2, 2, 1, -1, -2, -2};
  const int q_lo[8] = {2, 2, 1, -1, -2, -2, -1, 1};
  if (threadIdx.x < 8) {
    int k = threadIdx.x;
    lut[k].x = (uint32_t)(i_lo[k] + 65536 * q_lo[k]);
    lut[k].y = (uint32_t)(q_lo[k] - 65536 * i_lo[k]);
  }
}

__device__ __forceinline__ size_t bytes_for(int fmt, long long nsamples) {
  return fmt == GNSSB200_FMT_INT8_IQ ? (size_t)nsamples * 2 : (fmt == GNSSB200_FMT_PACKED2 ? (size_t)nsamples / 2 : (size_t)nsamples);
}

// one complex sample from a block, any format (slow, used by the serial path, tails, straddle fix-up)
__device__ __forceinline__ void load_sample(const uint8_t *blk, int fmt, int i, int &I, int &Q) {
  if (fmt == GNSSB200_FMT_INT8_IQ) {
    const int8_t *p = (const int8_t *)blk + 2 * (size_t)i;
    I = p[0];
    Q = p[1];
  } else if (fmt == GNSSB200_FMT_PACKED2) {
    uint32_t b = blk[i >> 1] >> ((i & 1) * 4);
    const int val[4] = {1, -1, 3, -3};  // FE/.../win32_sampler.h:45-55
    I = val[b & 3];
    Q = val[(b >> 2) & 3];
  } else {
    I = ((const int8_t *)blk)[i];
    Q = 0;
  }
}

// SPT consecutive samples starting at i0 -> SPT/2 words of (I0,Q0,I1,Q1) int8
// One packed byte (I0 Q0 I1 Q1 as 2-bit codes, LSB first) -> int8 word (I0,Q0,I1,Q1).  The four codes are
// spread into selector nibbles and one PRMT picks the values {+1,-1,+3,-3} from a register table:
// no shared-memory look-up, no bank conflicts.
__device__ __forceinline__ uint32_t unpack_byte(uint32_t b) {
  uint32_t x = (b | (b << 4)) & 0x0F0Fu;
  x = (x | (x << 2)) & 0x3333u;
  return __byte_perm(0xFD03FF01u, 0u, x);  // bytes: code0 -> 0x01, code1 -> 0xFF, code2 -> 0x03, code3 -> 0xFD
}

template <bool SMEM, class T>
__device__ __forceinline__ T ld_in(const T *p) {
  if constexpr (SMEM)
    return *p;  // shared-memory tile written by the TMA engine
  else
    return __ldg(p);
}

template <int SPT, bool SMEM = false>
__device__ __forceinline__ void load_chunk(const uint8_t *blk, int fmt, int i0, int nsamp, bool aligned,
                                           const uint32_t *unpack_lut, uint32_t (&w)[SPT / 2]) {
  if (i0 + SPT <= nsamp && aligned) {
    if (fmt == GNSSB200_FMT_INT8_IQ) {
      const uint4 *p = reinterpret_cast<const uint4 *>(blk + 2 * (size_t)i0);
#pragma unroll
      for (int q = 0; q < SPT / 8; q++) {
        uint4 v = ld_in<SMEM>(p + q);
        w[4 * q + 0] = v.x;
        w[4 * q + 1] = v.y;
        w[4 * q + 2] = v.z;
        w[4 * q + 3] = v.w;
      }
    } else if (fmt == GNSSB200_FMT_PACKED2) {
      // SPT/2 bytes; each byte -> one word through the 256-entry shared LUT
      const uint32_t *p = reinterpret_cast<const uint32_t *>(blk + (size_t)(i0 >> 1));
#pragma unroll
      for (int q = 0; q < SPT / 8; q++) {
        uint32_t v = ld_in<SMEM>(p + q);
        w[4 * q + 0] = unpack_byte(v & 0xff);
        w[4 * q + 1] = unpack_byte((v >> 8) & 0xff);
        w[4 * q + 2] = unpack_byte((v >> 16) & 0xff);
        w[4 * q + 3] = unpack_byte(v >> 24);
      }
    } else {
      const uint32_t *p = reinterpret_cast<const uint32_t *>(blk + (size_t)i0);
#pragma unroll
      for (int q = 0; q < SPT / 4; q++) {
        uint32_t v = ld_in<SMEM>(p + q);
        w[2 * q + 0] = __byte_perm(v, 0, 0x4140);  // (I0,0,I1,0)
        w[2 * q + 1] = __byte_perm(v, 0, 0x4342);
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < SPT / 2; q++) {
      uint32_t word = 0;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int i = i0 + 2 * q + e;
        if (i < nsamp) {
          int I, Q;
          load_sample(blk, fmt, i, I, Q);
          word |= ((uint32_t)(I & 0xff) | ((uint32_t)(Q & 0xff) << 8)) << (16 * e);
        }
      }
      w[q] = word;
    }
  }
}

// ---- TMA / mbarrier helpers (sm_90+ PTX) ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  // a plain test first: in the steady state the phase has completed long ago, and test_wait answers
  // faster than try_wait (which may suspend the thread for a hardware time slice)
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// The per-thread hot loop: SPT samples in groups of four.  Requires 1 <= kinc and 4*kinc < 2^32
// (checked by prepare_block), so at most one code-NCO carry falls inside a group: samples before it
// use the current E/P/L bits, samples after it the next table entry (correlator.c:227-250).
template <int SPT>
__device__ __forceinline__ void correlate_chunk(const uint32_t (&w)[SPT / 2], uint32_t cph, uint32_t kph,
                                                const uint32_t cinc, const uint32_t kinc, const uint32_t *tbl,
                                                uint32_t h, uint32_t bits, const uint2 *lut, int &accE, int &accP,
                                                int &accL) {
  int oE = sext8(bits, 0), oP = sext8(bits, 1), oL = sext8(bits, 2);
  int aE = 0, aP = 0, aL = 0;
  uint32_t hp = smem_u32(tbl + h);
  const uint32_t t1 = 0u - kinc, t2 = 0u - 2u * kinc, t3 = 0u - 3u * kinc, k4 = 4u * kinc;
#pragma unroll
  for (int g8 = 0; g8 < SPT / 8; g8++) {
    // eight LO look-ups in flight before the first product (shared-memory latency ~30 cycles)
    uint2 ab[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      ab[j] = lut[cph >> 29];
      cph += cinc;
    }
    int v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t word = w[(8 * g8 + j) >> 1];
      const int I = sext8(word, (j & 1) * 2), Q = sext8(word, (j & 1) * 2 + 1);
      v[j] = I * (int)ab[j].x + Q * (int)ab[j].y;  // ival + 65536*qval
    }
#pragma unroll
    for (int g = 0; g < 2; g++) {
      // sample j >= 1 still sees the old bits iff no carry happened in samples 0..j-1: kph + j*kinc < 2^32
      const bool m1 = kph < t1, m2 = kph < t2, m3 = kph < t3;
      int so = v[4 * g], sn = 0;
      if (m1) so += v[4 * g + 1]; else sn += v[4 * g + 1];
      if (m2) so += v[4 * g + 2]; else sn += v[4 * g + 2];
      if (m3) so += v[4 * g + 3]; else sn += v[4 * g + 3];
      uint32_t carry;
      asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(kph), "=r"(carry) : "r"(k4));
      hp += carry << 2;
      uint32_t t;
      asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(hp));
      const int nE = sext8(t, 0), nP = sext8(t, 1), nL = sext8(t, 2);
      aE += oE * so + nE * sn;
      aP += oP * so + nP * sn;
      aL += oL * so + nL * sn;
      oE = nE;
      oP = nP;
      oL = nL;
    }
  }
  accE = aE;
  accP = aP;
  accL = aL;
}

// Packed-native variant of the hot loop (GNSSB200_FMT_PACKED2, 32 samples = 16 bytes per thread): the
// 4-bit sample code (I,Q) and the 3-bit LO phase index a table of the finished mixer outputs
//   vlut[phase*16 + code][lane] = I*A[phase] + Q*B[phase]   (= ival + 65536*qval, exact small integers),
// replicated per lane so the look-up is bank-conflict free.  No unpack, no multiplies in the mixer.
struct PipeK {
  uint32_t k1, k8, k128, k2048;
};
template <int SPT>
__device__ __forceinline__ void correlate_chunk_packed(const uint32_t (&p)[SPT / 8], uint32_t cph, uint32_t kph,
                                                       const uint32_t cinc, const uint32_t kinc, const uint32_t *tbl,
                                                       uint32_t h, uint32_t bits, const uint32_t vlut_lane /* smem byte address of vlut[0][lane] */,
                                                       const PipeK K, int &accE, int &accP, int &accL) {
  int oE = sext8(bits, 0), oP = sext8(bits, 1), oL = sext8(bits, 2);
  int aE = 0, aP = 0, aL = 0;
  uint32_t hp = smem_u32(tbl + h);
  const uint32_t t1 = 0u - kinc, t2 = 0u - 2u * kinc, t3 = 0u - 3u * kinc, k4 = 4u * kinc;
#pragma unroll
  for (int g8 = 0; g8 < SPT / 8; g8++) {
    const uint32_t word = p[g8];
    // the eight 4-bit sample codes of this word as bytes: even samples in we, odd samples in wo
    const uint32_t we = word & 0x0F0F0F0Fu, wo = (word >> 4) & 0x0F0F0F0Fu;
    int v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      // entry offset = (phase*16 + code) * 128 bytes.  One PRMT on the ALU pipe (the code byte); the LO phase
      // (top three bits of the carrier NCO), both scalings and the NCO step are IMADs on the FMA pipe.
      const uint32_t code = __byte_perm((j & 1) ? wo : we, 0u, 0x4440u | (uint32_t)(j >> 1));
      const uint32_t idx = __umulhi(cph, K.k8);
      const uint32_t addr = idx * K.k2048 + (code * K.k128 + vlut_lane);
      uint32_t t;
      asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(addr));
      v[j] = (int)t;
      cph = cinc * K.k1 + cph;
    }
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const bool m1 = kph < t1, m2 = kph < t2, m3 = kph < t3;
      int so = v[4 * g], sn = 0;
      if (m1) so += v[4 * g + 1]; else sn += v[4 * g + 1];
      if (m2) so += v[4 * g + 2]; else sn += v[4 * g + 2];
      if (m3) so += v[4 * g + 3]; else sn += v[4 * g + 3];
      uint32_t carry;
      asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(kph), "=r"(carry) : "r"(k4));
      hp += carry << 2;
      uint32_t t;
      asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(hp));
      const int nE = sext8(t, 0), nP = sext8(t, 1), nL = sext8(t, 2);
      aE += oE * so + nE * sn;
      aP += oP * so + nP * sn;
      aL += oL * so + nL * sn;
      oE = nE;
      oP = nP;
      oL = nL;
    }
  }
  accE = aE;
  accP = aP;
  accL = aL;
}

__device__ __forceinline__ void unpack_lanes(int packed, int &lo, int &hi) {
  lo = (int)(short)(packed & 0xffff);
  hi = (packed - lo) >> 16;
}

// ---- lane-0 bookkeeping -----------------------------------------------------------------------
struct ChanShared {
  gnssb200_chan k;
  gnssb200_corr g;
  ChRegs r;
  long long tic;
  int dumped_last;
  int halted;
  int dump_count;
};

// epoch-counter load request of the host logic, applied at the start of a block (correlator.c:177-182)
__device__ __forceinline__ void apply_epoch_load(ChanShared &cs) {
  ChRegs &r = cs.r;
  if (r.w_epoch != -1) {
    r.r_meas[7] = r.w_epoch;
    cs.g.ms_counter = r.w_epoch & 0xff;
    cs.g.bit_counter = r.w_epoch >> 8;
    r.w_epoch = -1;
  }
}

// correlator parameters of the next block, part 1: what the correlator state alone decides (TIC
// down-counter, NCO phases, half-chip count) -- independent of the channel's write registers
__device__ __forceinline__ void prepare_block_state(ChanShared &cs, StepParams &sp, const TrackArgs &a) {
  const long long n = a.nsamp;
  if (cs.tic < n) {  // correlator.c:155-165
    sp.tic_count = (int)cs.tic;
    cs.tic += a.cfg.tic_ref - n;
  } else {
    cs.tic -= n;
    sp.tic_count = -1;
  }
  sp.tic = cs.tic;
  sp.cyc_pending = 0;
  sp.cph0 = cs.g.carrier_phase;
  sp.kph0 = cs.g.code_phase;
  sp.hc0 = cs.g.half_chip & 0xffff;
}
// part 2: NCO increments, dump position and path selection from the write registers
__device__ __forceinline__ void prepare_block_regs(ChanShared &cs, StepParams &sp, const TrackArgs &a, int tbl_prn) {
  const long long n = a.nsamp;
  ChRegs &r = cs.r;
  if (r.w_prn <= 0) {
    sp.mode = MODE_IDLE;
    return;
  }
  sp.cinc = (uint32_t)((r.w_carr_hi << 16) + r.w_carr_lo);
  sp.kinc = (uint32_t)((r.w_code_hi << 16) + r.w_code_lo) << 1;
  const long long slew_dump = (long long)r.w_slew + HALF_CHIPS;  // :172
  sp.slew_dump = (uint32_t)slew_dump;
  const long long w1 = ((long long)sp.hc0 + 1 >= slew_dump) ? 1 : slew_dump - sp.hc0;
  const unsigned long long wtot = ((unsigned long long)sp.kph0 + (unsigned long long)n * sp.kinc) >> 32;
  sp.w1 = (uint32_t)w1;
  sp.stale_idx = (uint32_t)(sp.hc0 + w1);
  bool fast = (r.w_prn == tbl_prn) && r.w_prn >= 1 && r.w_prn <= 32 && slew_dump >= 1 && slew_dump < 65536 &&
              (long long)wtot < w1 + slew_dump && (wtot + 40) < SMEM_TBL && (sp.hc0 + w1 + 40) < SMEM_TBL &&
              n < (1ll << 30) && sp.kinc >= 1u && sp.kinc < (1u << 30);
  sp.mode = fast ? MODE_FAST : MODE_SERIAL;
}
__device__ __forceinline__ void prepare_block_params(ChanShared &cs, StepParams &sp, const TrackArgs &a, int tbl_prn) {
  prepare_block_state(cs, sp, a);
  prepare_block_regs(cs, sp, a, tbl_prn);
}

__device__ __forceinline__ void prepare_block(ChanShared &cs, StepParams &sp, const TrackArgs &a, int tbl_prn) {
  apply_epoch_load(cs);
  prepare_block_params(cs, sp, a, tbl_prn);
}

// dump side effects common to both paths (correlator.c:252-281)
__device__ __forceinline__ void apply_dump_counters(ChanShared &cs) {
  gnssb200_corr &g = cs.g;
  cs.r.w_slew = 0;
  g.ms_counter++;
  if (g.ms_counter == 20) g.bit_counter = (g.bit_counter + 1) % 50;
  g.ms_counter %= 20;
  cs.r.r_meas[7] = g.ms_counter + (g.bit_counter << 8);
}

// End-of-block rules of the closed-form path, split in two: everything that follows from the block's
// parameters alone (dump or not, counters, half-chip count, TIC latch, NCO phases) ...
__device__ __forceinline__ void finalize_state(ChanShared &cs, const StepParams &sp, int nsamp) {
  gnssb200_corr &g = cs.g;
  ChRegs &r = cs.r;
  const unsigned long long n = (unsigned long long)nsamp;
  const unsigned long long kend = (unsigned long long)sp.kph0 + n * sp.kinc;
  const unsigned long long cend = (unsigned long long)sp.cph0 + n * sp.cinc;
  const uint32_t wtot = (uint32_t)(kend >> 32);
  const bool dumped = sp.w1 <= wtot;
  const int epoch_before = r.r_meas[7];
  if (dumped) {
    apply_dump_counters(cs);
    g.half_chip = wtot - sp.w1;
  } else {
    g.half_chip = sp.hc0 + wtot;
  }
  cs.dumped_last = dumped ? 1 : 0;
  const uint32_t ctot = (uint32_t)(cend >> 32);
  if (sp.tic_count >= 0 && sp.tic_count < nsamp) {  // measurement latch, :286-303
    const unsigned long long m = (unsigned long long)sp.tic_count + 1;
    const unsigned long long kt = (unsigned long long)sp.kph0 + m * sp.kinc;
    const unsigned long long ct = (unsigned long long)sp.cph0 + m * sp.cinc;
    const uint32_t wa = (uint32_t)(kt >> 32), cw = (uint32_t)(ct >> 32);
    const bool by_tic = dumped && sp.w1 <= wa;
    r.r_meas[4] = by_tic ? r.r_meas[7] : epoch_before;
    r.r_meas[3] = (int)((uint32_t)ct >> 22);
    r.r_meas[1] = (int)(by_tic ? wa - sp.w1 : sp.hc0 + wa);
    r.r_meas[5] = (int)((uint32_t)kt >> 22);
    const uint32_t cyc = g.carrier_cycle + cw;
    r.r_meas[2] = (int)(cyc & 0xffff);
    r.r_meas[6] = (int)(cyc >> 16);
    g.carrier_cycle = ctot - cw;
  } else {
    g.carrier_cycle += ctot;
  }
  g.carrier_phase = (uint32_t)cend;
  g.code_phase = (uint32_t)kend;
}
// ... and the accumulators, which need the block's sums: A = samples up to the dump (or all of them),
// B = samples after it.  Requires cs.dumped_last from finalize_state.
__device__ __forceinline__ void finalize_acc(ChanShared &cs, const int (&A)[6], const int (&B)[6]) {
  gnssb200_corr &g = cs.g;
  if (cs.dumped_last) {
#pragma unroll
    for (int q = 0; q < 6; q++) {
      cs.r.r_acc[q] = g.acc[q] + A[q];
      g.acc[q] = B[q];
    }
  } else {
#pragma unroll
    for (int q = 0; q < 6; q++) g.acc[q] += A[q] + B[q];
  }
}
__device__ __forceinline__ void finalize_fast(ChanShared &cs, const StepParams &sp, const int (&A)[6], const int (&B)[6], int nsamp) {
  finalize_state(cs, sp, nsamp);
  finalize_acc(cs, A, B);
}

// literal per-sample walk of one block by a single lane (any register contents)
__device__ __noinline__ void serial_block(ChanShared &cs, const StepParams &sp, const uint32_t *code_table, int fmt, int nsamp,
                                          const uint8_t *blk) {
  const int i_lo[8] = {-1, 1, 2, 2, 1, -1, -2, -2};
  const int q_lo[8] = {2, 2, 1, -1, -2, -2, -1, 1};
  gnssb200_corr &g = cs.g;
  ChRegs &r = cs.r;
  const long long row = (long long)r.w_prn * HALF_CHIPS;
  const int dump_at = r.w_slew + HALF_CHIPS;
  uint16_t hc = (uint16_t)g.half_chip;
  auto bits_at = [&](uint16_t hh) -> uint32_t {
    long long f = row + hh;
    return (f >= 0 && f < TABLE_ENTRIES) ? code_table[f] : 0u;
  };
  uint32_t t = bits_at(hc);
  int cE = sext8(t, 0), cP = sext8(t, 1), cL = sext8(t, 2);
  int dumped = 0;
  for (int i = 0; i < nsamp; i++) {
    const int k = g.carrier_phase >> 29;
    int I, Q;
    load_sample(blk, fmt, i, I, Q);
    const int vq = q_lo[k] * I - i_lo[k] * Q;
    const int vi = i_lo[k] * I + q_lo[k] * Q;
    g.acc[0] += cL * vi;
    g.acc[1] += cL * vq;
    g.acc[2] += cP * vi;
    g.acc[3] += cP * vq;
    g.acc[4] += cE * vi;
    g.acc[5] += cE * vq;
    uint32_t before = g.carrier_phase;
    g.carrier_phase += sp.cinc;
    if (g.carrier_phase < before) g.carrier_cycle++;
    before = g.code_phase;
    g.code_phase += sp.kinc;
    if (g.code_phase < before) {
      hc++;
      t = bits_at(hc);
      cE = sext8(t, 0);
      cP = sext8(t, 1);
      cL = sext8(t, 2);
      if (hc >= dump_at) {
        for (int q = 0; q < 6; q++) {
          r.r_acc[q] = g.acc[q];
          g.acc[q] = 0;
        }
        apply_dump_counters(cs);
        hc = 0;
        dumped = 1;
      }
    }
    if (i == sp.tic_count) {
      r.r_meas[4] = r.r_meas[7];
      r.r_meas[3] = (int)(g.carrier_phase >> 22);
      r.r_meas[1] = hc;
      r.r_meas[5] = (int)(g.code_phase >> 22);
      r.r_meas[2] = (int)(g.carrier_cycle & 0xffff);
      r.r_meas[6] = (int)(g.carrier_cycle >> 16);
      g.carrier_cycle = 0;
    }
  }
  g.half_chip = hc;
  cs.dumped_last = dumped;
}

__device__ __forceinline__ void after_block(ChanShared &cs, const TrackArgs &a, int s, int ch, long long block_index) {
  if (!cs.dumped_last) return;
  if (a.run_isr) {
    if (dev_gpsisr_channel(cs.k, cs.r, a.cfg)) {
      cs.halted = 1;
      return;
    }
  }
  if (a.dumps && cs.dump_count < a.dump_cap) {
    // 48-byte record written as three 16-byte stores
    gnssb200_dump *out = &a.dumps[((size_t)s * NCH + ch) * a.dump_cap + cs.dump_count];
    int4 q0, q1, q2;
    q0.x = (int)block_index;
    q0.y = (int)(uint16_t)(int16_t)ch | ((int)(uint16_t)(int16_t)cs.k.state << 16);
    q0.z = cs.r.r_acc[0];
    q0.w = cs.r.r_acc[1];
    q1.x = cs.r.r_acc[2];
    q1.y = cs.r.r_acc[3];
    q1.z = cs.r.r_acc[4];
    q1.w = cs.r.r_acc[5];
    q2.x = (cs.r.w_carr_hi << 16) + cs.r.w_carr_lo;
    q2.y = (cs.r.w_code_hi << 16) + cs.r.w_code_lo;
    q2.z = (int)(uint16_t)(int16_t)cs.k.n_freq | ((int)(uint16_t)(int16_t)cs.k.codes << 16);
    q2.w = cs.r.w_slew;
    int4 *o4 = reinterpret_cast<int4 *>(out);
    o4[0] = q0;
    o4[1] = q1;
    o4[2] = q2;
    cs.dump_count++;
  }
}

// sum over the warp of six ints (butterfly); result valid in every lane
__device__ __forceinline__ void warp_sum6(int (&v)[6]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 6; q++) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
}

// dynamic shared memory: two sample tiles of tile_bytes each (TMA mode only)
// FMT >= 0 fixes the sample format at compile time (the TMA-staged hot variants); FMT < 0 reads it from
// the arguments (generic variant: unaligned / ragged / I-only blocks, loaded straight from global memory).
template <int MAXT, int MINB, int FMT, bool TMA, int SPT = 32>
__global__ void __launch_bounds__(MAXT, MINB) track_loop_kernel(const TrackArgs a, const int tile_bytes) {
  constexpr bool use_tma = TMA;
  const int fmt = FMT >= 0 ? FMT : a.fmt;
  __shared__ ChanShared cs;
  __shared__ StepParams sp_s;
  __shared__ uint2 lut[8];
  __shared__ uint32_t tbl[SMEM_TBL];
  // copy of tbl[0..47] whose entry 0 holds the bits left over from the dump (rule A6): a chunk that
  // starts in the first post-dump half chip walks alias_tbl[0] -> tbl[1] -> tbl[2] ... like the reference
  __shared__ uint32_t alias_tbl[48];
  __shared__ uint32_t unpack_lut[256];
  __shared__ __align__(16) int totals[12];  // six pre-dump and six post-dump sums, accumulated by shared-memory atomics
  __shared__ __align__(8) uint64_t mbar[2];
  extern __shared__ __align__(128) uint8_t tiles[];
  constexpr bool packed_native = TMA && FMT == GNSSB200_FMT_PACKED2;
  // [128 entries][32 lanes] mixer-output table of the packed-native loop, placed after the two tiles
  uint32_t *vlut = reinterpret_cast<uint32_t *>(tiles + 2 * (size_t)tile_bytes);

  const int s = a.first_stream + blockIdx.x / NCH, ch = blockIdx.x % NCH;
  gnssb200_rx *rx = a.rx + s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tbl_prn = rx->reg_write[ch << 3];

  fill_lo_lut(lut);
  if (tid < 12) totals[tid] = 0;
  for (int i = tid; i < SMEM_TBL; i += blockDim.x) {
    long long f = (long long)tbl_prn * HALF_CHIPS + i;
    tbl[i] = (tbl_prn >= 1 && tbl_prn <= 32 && f < TABLE_ENTRIES) ? a.code_table[f] : 0u;
  }
  if (packed_native) {
    const int i_lo[8] = {-1, 1, 2, 2, 1, -1, -2, -2};
    const int q_lo[8] = {2, 2, 1, -1, -2, -2, -1, 1};
    const int val[4] = {1, -1, 3, -3};
    for (int i = tid; i < 128 * 32; i += blockDim.x) {
      const int e = i >> 5, ph = e >> 4, code = e & 15;
      const int I = val[code & 3], Q = val[code >> 2];
      const int ival = i_lo[ph] * I + q_lo[ph] * Q, qval = q_lo[ph] * I - i_lo[ph] * Q;  // correlator.c:214-215
      vlut[i] = (uint32_t)(ival + 65536 * qval);
    }
  }
  if (tid < 48) {
    long long f = (long long)tbl_prn * HALF_CHIPS + tid;
    alias_tbl[tid] = (tbl_prn >= 1 && tbl_prn <= 32 && f < TABLE_ENTRIES) ? a.code_table[f] : 0u;
  }
  for (int i = tid; i < 256; i += blockDim.x) {
    const int val[4] = {1, -1, 3, -3};
    uint32_t wv = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) wv |= (uint32_t)(val[(i >> (2 * e)) & 3] & 0xff) << (8 * e);
    unpack_lut[i] = wv;
  }
  const size_t blk_bytes = bytes_for(fmt, a.nsamp);
  const uint8_t *stream_base = a.d_if + (size_t)s * a.stride;
  const bool aligned = ((reinterpret_cast<uintptr_t>(stream_base) | blk_bytes) & 15) == 0 && (a.nsamp % 8) == 0;
  if (tid == 0) {
    cs.k = rx->chan[ch];
    cs.g = rx->corr[ch];
    const int b8 = ch << 3;
    cs.r.w_prn = rx->reg_write[b8];
    cs.r.w_carr_hi = rx->reg_write[b8 + 3];
    cs.r.w_carr_lo = rx->reg_write[b8 + 4];
    cs.r.w_code_hi = rx->reg_write[b8 + 5];
    cs.r.w_code_lo = rx->reg_write[b8 + 6];
    cs.r.w_epoch = rx->reg_write[b8 + 7];
    cs.r.w_slew = rx->reg_write[b8 + 0x84];
    for (int q = 0; q < 8; q++) cs.r.r_meas[q] = rx->reg_read[b8 + q];
    for (int q = 0; q < 6; q++) cs.r.r_acc[q] = rx->reg_read[b8 + 0x84 + q];
    cs.tic = rx->tic;
    cs.dumped_last = 0;
    cs.halted = 0;
    cs.dump_count = a.dump_count ? a.dump_count[s * NCH + ch] : 0;
    if (a.nblocks > 0 && !rx->halted)
      prepare_block(cs, sp_s, a, tbl_prn);
    else
      sp_s.mode = MODE_STOP;
    sp_s.stale_bits = 0;
    if (sp_s.mode == MODE_FAST) {
      const long long f = (long long)tbl_prn * HALF_CHIPS + sp_s.stale_idx;
      sp_s.stale_bits = f < TABLE_ENTRIES ? a.code_table[f] : 0u;
    }
    alias_tbl[0] = sp_s.stale_bits;
    if (use_tma) {
      mbar_init(&mbar[0], 1);
      mbar_init(&mbar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (sp_s.mode != MODE_STOP && sp_s.mode != MODE_IDLE) {  // block 0 -> tile 0
        mbar_expect_tx(&mbar[0], (uint32_t)blk_bytes);
        tma_load_1d(tiles, stream_base, (uint32_t)blk_bytes, &mbar[0]);
      }
    }
  }
  __syncthreads();
  const long long first_block = rx->blocks_done;
  const int my_i0 = tid * SPT;

#ifdef TRACK_PROFILE
  long long t_main = 0, t_red = 0, t_isr = 0, t_sync2 = 0, t_corr = 0, t_quiet = 0, n_quiet = 0, t_fin = 0, t_after = 0, t_prep = 0, t_wait = 0, t_head = 0, t_load = 0, t_setup = 0, t_post = 0;
  long long t_state[8] = {0,0,0,0,0,0,0,0}, n_state[8] = {0,0,0,0,0,0,0,0};
#endif
  // Every thread keeps a (uniform) copy of the block parameters.  A block without dump, TIC latch or
  // mode change is "quiet": nothing leaves the thread -- its sums are carried in registers, the
  // parameters of the next block follow from the closed forms, and neither the reduction nor the
  // ISR lane runs.  Only blocks with an event (a dump, about every second block) synchronise.
  StepParams sp = sp_s;
  int carry[6] = {0, 0, 0, 0, 0, 0};
  unsigned long long nk = (unsigned long long)a.nsamp * sp.kinc, nc = (unsigned long long)a.nsamp * sp.cinc;
  for (long long b = 0; b < a.nblocks; b++) {
#ifdef TRACK_PROFILE
    long long c0 = clock64();
#endif
    if (sp.mode == MODE_STOP) break;
    const uint8_t *blk = stream_base + (size_t)b * blk_bytes;
    int sumA[6] = {0, 0, 0, 0, 0, 0}, sumB[6] = {0, 0, 0, 0, 0, 0};
    bool anyB = false;
    // parameters the next block would have if this one is quiet
    StepParams nx = sp;
    bool quiet = false;
    if (sp.mode == MODE_FAST && b + 1 < a.nblocks) {
      const unsigned long long n = (unsigned long long)a.nsamp;
      const unsigned long long kend = (unsigned long long)sp.kph0 + nk;
      const unsigned long long cend = (unsigned long long)sp.cph0 + nc;
      const uint32_t wtot = (uint32_t)(kend >> 32);
      nx.kph0 = (uint32_t)kend;
      nx.cph0 = (uint32_t)cend;
      nx.hc0 = sp.hc0 + wtot;
      nx.w1 = sp.w1 - wtot;
      nx.cyc_pending = sp.cyc_pending + (uint32_t)(cend >> 32);
      if (sp.tic < (long long)n) {
        nx.tic_count = (int)sp.tic;
        nx.tic = sp.tic + a.cfg.tic_ref - (long long)n;
      } else {
        nx.tic_count = -1;
        nx.tic = sp.tic - (long long)n;
      }
      const unsigned long long wnext = ((unsigned long long)nx.kph0 + nk) >> 32;
      const bool next_fast = wnext < (unsigned long long)nx.w1 + sp.slew_dump && (wnext + 40) < SMEM_TBL;
      quiet = wtot < sp.w1 && !(sp.tic_count >= 0 && sp.tic_count < a.nsamp) && next_fast;
    }
    const uint8_t *tile = tiles + (size_t)(b & 1) * tile_bytes;

    if (use_tma && sp.mode != MODE_IDLE) {
      // prefetch block b+1 into the other tile (every thread finished reading it before the barrier
      // that ended block b-1), then wait for block b
      if (tid == 0 && b + 1 < a.nblocks) {
        mbar_expect_tx(&mbar[(b + 1) & 1], (uint32_t)blk_bytes);
        tma_load_1d(tiles + (size_t)((b + 1) & 1) * tile_bytes, blk + blk_bytes, (uint32_t)blk_bytes, &mbar[(b + 1) & 1]);
      }
#ifdef TRACK_PROFILE
      long long w0c = clock64();
#endif
      mbar_wait(&mbar[b & 1], (uint32_t)((b >> 1) & 1));
#ifdef TRACK_PROFILE
      t_wait += clock64() - w0c;
#endif
    }
#ifdef TRACK_PROFILE
    t_head += clock64() - c0;
#endif

    if (sp.mode == MODE_FAST) {
      // trip count is uniform over the CTA (warp collectives inside); `live` masks ragged tails
      for (int base = 0; base < a.nsamp; base += blockDim.x * SPT) {
        const int i0 = base + my_i0;
        const bool live = i0 < a.nsamp;
        uint32_t cur[SPT / 2];
        uint32_t pk[SPT / 8];
#ifdef TRACK_PROFILE
        long long l0c = clock64();
#endif
        if (live && packed_native) {
          const uint32_t *pp = reinterpret_cast<const uint32_t *>(tile + (i0 >> 1));
#pragma unroll
          for (int q = 0; q < SPT / 8; q++) pk[q] = pp[q];
        } else if (live) {
          if (use_tma)
            load_chunk<SPT, true>(tile, fmt, i0, a.nsamp, true, unpack_lut, cur);  // shared-memory tile
          else
            load_chunk<SPT>(blk, fmt, i0, a.nsamp, aligned, unpack_lut, cur);
        }
#ifdef TRACK_PROFILE
        t_load += clock64() - l0c;
#endif
        const int i1 = live ? min(i0 + SPT, a.nsamp) : i0 + 1;
        const unsigned long long k0 = (unsigned long long)sp.kph0 + (unsigned long long)i0 * sp.kinc;
        const uint32_t w_start = (uint32_t)(k0 >> 32);
        const uint32_t w_lastb = (uint32_t)(((unsigned long long)sp.kph0 + (unsigned long long)(i1 - 1) * sp.kinc) >> 32);
        const bool allA = !live || w_lastb < sp.w1, allB = live && w_start >= sp.w1;
        uint32_t h, hl;
        if (allB) {
          h = w_start - sp.w1;
          hl = (h == 0) ? sp.stale_idx : h;  // stale bits after the dump (SURVEY.md App. A rule A6)
        } else {
          h = sp.hc0 + w_start;
          hl = h;
        }
        int pE = 0, pP = 0, pL = 0;
#ifdef TRACK_PROFILE
        long long cc0 = clock64();
#endif
        // chunk starting in the first post-dump half chip: stale bits first, then tbl[1], tbl[2], ...
        const bool stale_start = allB && h == 0;
        if (live && packed_native)
          correlate_chunk_packed<SPT>(pk, sp.cph0 + (uint32_t)i0 * sp.cinc, (uint32_t)k0, sp.cinc, sp.kinc,
                                      stale_start ? alias_tbl : tbl, h, stale_start ? sp.stale_bits : tbl[hl],
                                      smem_u32(vlut) + 4u * (uint32_t)lane, PipeK{a.k1, a.k8, a.k128, a.k2048}, pE, pP, pL);
        else if (live)
          correlate_chunk<SPT>(cur, sp.cph0 + (uint32_t)i0 * sp.cinc, (uint32_t)k0, sp.cinc, sp.kinc,
                               stale_start ? alias_tbl : tbl, h, stale_start ? sp.stale_bits : tbl[hl], lut, pE, pP, pL);
#ifdef TRACK_PROFILE
        t_corr += clock64() - cc0 + (pE & 0);
        t_setup += cc0 - l0c;
        long long p0c = clock64();
#endif
        const bool straddle = !allA && !allB;
        if (!straddle && live) {
          int v[6];
          unpack_lanes(pL, v[0], v[1]);
          unpack_lanes(pP, v[2], v[3]);
          unpack_lanes(pE, v[4], v[5]);
          if (allA) {
#pragma unroll
            for (int q = 0; q < 6; q++) sumA[q] += v[q];
          } else {
#pragma unroll
            for (int q = 0; q < 6; q++) sumB[q] += v[q];
          }
        }
        // the chunk that contains the dump is re-evaluated one sample per lane by its warp
        unsigned m = __ballot_sync(0xffffffffu, straddle);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const int si0 = __shfl_sync(0xffffffffu, i0, src);
          for (int i = si0 + lane; i < min(si0 + SPT, a.nsamp); i += 32) {
            const unsigned long long ki = (unsigned long long)sp.kph0 + (unsigned long long)i * sp.kinc;
            const uint32_t wb = (uint32_t)(ki >> 32);
            const bool inA = wb < sp.w1;
            const uint32_t rel = wb - sp.w1;
            const uint32_t hh = inA ? sp.hc0 + wb : (rel == 0 ? sp.stale_idx : rel);
            const uint32_t t = tbl[hh];
            int I, Q;
            load_sample(use_tma ? tile : blk, fmt, i, I, Q);
            const uint2 ab = lut[(sp.cph0 + (uint32_t)i * sp.cinc) >> 29];
            const int v = I * (int)ab.x + Q * (int)ab.y;
            int vi, vq;
            unpack_lanes(v, vi, vq);
            const int cE = sext8(t, 0), cP = sext8(t, 1), cL = sext8(t, 2);
            if (inA) {
              sumA[0] += cL * vi; sumA[1] += cL * vq; sumA[2] += cP * vi;
              sumA[3] += cP * vq; sumA[4] += cE * vi; sumA[5] += cE * vq;
            } else {
              sumB[0] += cL * vi; sumB[1] += cL * vq; sumB[2] += cP * vi;
              sumB[3] += cP * vq; sumB[4] += cE * vi; sumB[5] += cE * vq;
            }
          }
        }
        anyB |= !allA;
#ifdef TRACK_PROFILE
        t_post += clock64() - p0c;
#endif
      }
      if (quiet) {  // no dump in this block: every chunk was pre-dump, keep the sums in registers
#pragma unroll
        for (int q = 0; q < 6; q++) carry[q] += sumA[q];
        sp = nx;
        __syncthreads();  // tile (b+1)&1 may be refilled by the TMA issue of the next iteration
#ifdef TRACK_PROFILE
        t_quiet += clock64() - c0;
        n_quiet++;
#endif
        continue;
      }
#pragma unroll
      for (int q = 0; q < 6; q++) {
        sumA[q] += carry[q];
        carry[q] = 0;
      }
      // warp reduction (shuffles); the post-dump set only where a warp has post-dump samples
      const bool warpB = __any_sync(0xffffffffu, anyB);
      warp_sum6(sumA);
      if (warpB) warp_sum6(sumB);
      if (lane < 6) {
        int va = sumA[0], vb = sumB[0];
#pragma unroll
        for (int q = 1; q < 6; q++) {
          if (lane == q) {
            va = sumA[q];
            vb = sumB[q];
          }
        }
        atomicAdd(&totals[lane], va);
        if (warpB) atomicAdd(&totals[6 + lane], vb);
      }
    }
#ifdef TRACK_PROFILE
    long long c1 = clock64();
#endif
    __syncthreads();
#ifdef TRACK_PROFILE
    long long c2 = clock64();
#endif

    if (warp == 0) {
      int A[6], B[6];
      if (sp.mode == MODE_FAST && tid == 0) {
        const int4 t0 = *reinterpret_cast<const int4 *>(&totals[0]);
        const int4 t1 = *reinterpret_cast<const int4 *>(&totals[4]);
        const int4 t2 = *reinterpret_cast<const int4 *>(&totals[8]);
        A[0] = t0.x; A[1] = t0.y; A[2] = t0.z; A[3] = t0.w; A[4] = t1.x; A[5] = t1.y;
        B[0] = t1.z; B[1] = t1.w; B[2] = t2.x; B[3] = t2.y; B[4] = t2.z; B[5] = t2.w;
        const int4 z = make_int4(0, 0, 0, 0);
        *reinterpret_cast<int4 *>(&totals[0]) = z;  // ready for the next event (barrier below orders it)
        *reinterpret_cast<int4 *>(&totals[4]) = z;
        *reinterpret_cast<int4 *>(&totals[8]) = z;
      }
      if (tid == 0) {
        // state that advanced in registers during quiet blocks
#ifdef TRACK_PROFILE
        long long i0c = clock64();
#endif
        cs.tic = sp.tic;
        cs.g.carrier_cycle += sp.cyc_pending;
        if (sp.mode == MODE_FAST)
          finalize_fast(cs, sp, A, B, a.nsamp);
        else if (sp.mode == MODE_SERIAL)
          serial_block(cs, sp, a.code_table, fmt, a.nsamp, use_tma ? tile : blk);
        else
          cs.dumped_last = 0;
#ifdef TRACK_PROFILE
        long long i1c = clock64();
#endif
        after_block(cs, a, s, ch, first_block + b);
#ifdef TRACK_PROFILE
        long long i2c = clock64();
        t_fin += i1c - i0c; t_after += i2c - i1c; t_state[cs.k.state & 7] += i2c - i1c; n_state[cs.k.state & 7]++;
#endif
        if (cs.halted || sp.mode == MODE_IDLE)  // an idle channel has no ISR: nothing can change any more
          sp_s.mode = MODE_STOP;
        else if (b + 1 < a.nblocks) {
          prepare_block(cs, sp_s, a, tbl_prn);
          if (sp_s.mode == MODE_FAST) {
            sp_s.stale_bits = tbl[sp_s.stale_idx];
            alias_tbl[0] = sp_s.stale_bits;  // other threads read it only after the barrier below
          }
        }
#ifdef TRACK_PROFILE
        t_prep += clock64() - i2c;
#endif
      }
    }
#ifdef TRACK_PROFILE
    long long c3 = clock64();
#endif
    __syncthreads();
    sp = sp_s;
    nk = (unsigned long long)a.nsamp * sp.kinc;
    nc = (unsigned long long)a.nsamp * sp.cinc;
#ifdef TRACK_PROFILE
    long long c4 = clock64();
    t_main += c1 - c0; t_red += c2 - c1; t_isr += c3 - c2; t_sync2 += c4 - c3;
#endif
  }
#ifdef TRACK_PROFILE
  if (blockIdx.x == 0 && (tid == 0 || tid == 37 || tid == 255))
  {
    const long long ne = a.nblocks - n_quiet;
    printf("tid %d: quiet blocks %lld x %lld cyc; event blocks %lld: main %lld sync1 %lld isr %lld sync2 %lld; corr/block %lld\n", tid, n_quiet,
           n_quiet ? t_quiet / n_quiet : 0, ne, t_main / ne, t_red / ne, t_isr / ne, t_sync2 / ne, t_corr / a.nblocks);
    printf("   tid %d: head(incl wait) %lld  mbar wait %lld  load %lld  load+setup %lld  post(unpack,straddle) %lld per block\n", tid, t_head / a.nblocks,
           t_wait / a.nblocks, t_load / a.nblocks, t_setup / a.nblocks, t_post / a.nblocks);
#ifdef TRACK_PROFILE_ISR
    if (tid == 0)
      printf("   isr sections (cycles per event block): primitives %lld  pll %lld  carrier word %lld  dll %lld  code word %lld  pull-in bookkeeping %lld\n",
             g_isr_t[0] / ne, g_isr_t[1] / ne, g_isr_t[2] / ne, g_isr_t[3] / ne, g_isr_t[4] / ne, g_isr_t[5] / ne);
#endif
    if (tid == 0)
      printf("   isr lane: finalize %lld  after_block %lld  prepare %lld per event; after_block by state after: acq %lld (%lld) conf %lld (%lld) pull %lld (%lld) track %lld (%lld)\n",
             t_fin / ne, t_after / ne, t_prep / ne, n_state[1] ? t_state[1] / n_state[1] : 0, n_state[1], n_state[2] ? t_state[2] / n_state[2] : 0, n_state[2],
             n_state[3] ? t_state[3] / n_state[3] : 0, n_state[3], n_state[4] ? t_state[4] / n_state[4] : 0, n_state[4]);
  }
#endif

  if (tid == 0) {
    rx->chan[ch] = cs.k;
    rx->corr[ch] = cs.g;
    const int b8 = ch << 3;
    rx->reg_write[b8 + 3] = cs.r.w_carr_hi;
    rx->reg_write[b8 + 4] = cs.r.w_carr_lo;
    rx->reg_write[b8 + 5] = cs.r.w_code_hi;
    rx->reg_write[b8 + 6] = cs.r.w_code_lo;
    rx->reg_write[b8 + 7] = cs.r.w_epoch;
    rx->reg_write[b8 + 0x84] = cs.r.w_slew;
    for (int q = 1; q < 8; q++) rx->reg_read[b8 + q] = cs.r.r_meas[q];
    for (int q = 0; q < 6; q++) rx->reg_read[b8 + 0x84 + q] = cs.r.r_acc[q];
    a.chan_flags[s * NCH + ch] = (cs.dumped_last ? 1 : 0) | (cs.halted ? 2 : 0);
    if (a.dump_count) a.dump_count[s * NCH + ch] = cs.dump_count;
  }
}

// ==================================================================================================
// Warp-specialised variant of the channel loop (the hot path for 8192-sample TMA-staged blocks).
//
// Eight correlator warps and one control lane per (stream, channel) CTA, decoupled by mbarriers:
//
//   control lane    issues the TMA load of block b+1, derives the parameters of block b+1 (closed
//                   forms for quiet blocks; reduction totals -> dump rules -> channel state machine
//                   for event blocks) and publishes them in a two-slot ring; the bookkeeping part of
//                   the ISR (bit sync, confirm counters, the dump record) runs AFTER the publish, while
//                   the correlator warps already work on the next block.
//   correlator warp waits for parameters + samples of block b, correlates its 8 x 32 x 32 samples,
//                   carries its sums in registers over quiet blocks; in an event block it reduces
//                   (REDUX.SUM) into shared-memory totals and signals the control lane.  No CTA-wide
//                   barrier in the loop; a warp is at most one block ahead of the slowest one.
//
// Same arithmetic, same rules, same results as track_loop_kernel (which remains the generic variant).
struct __align__(16) BlockParams {
  uint32_t cph0, kph0, cinc, kinc;
  uint32_t hc0, w1, stale_idx, stale_bits;
  int mode, event, pad0, pad1;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

// Is block `sp` an event block (reduce + control-lane work at its end)?  Same rule as the quiet test of
// track_loop_kernel: quiet = no dump, no TIC latch, not the last block, and the next block still fits the
// closed-form path.
__device__ __forceinline__ bool block_is_event(const StepParams &sp, const TrackArgs &a, bool last) {
  if (sp.mode != MODE_FAST || last) return true;
  const unsigned long long nk = (unsigned long long)a.nsamp * sp.kinc;
  const unsigned long long kend = (unsigned long long)sp.kph0 + nk;
  const uint32_t wtot = (uint32_t)(kend >> 32);
  const unsigned long long wnext = ((unsigned long long)(uint32_t)kend + nk) >> 32;
  const bool next_fast = wnext < (unsigned long long)(sp.w1 - wtot) + sp.slew_dump && (wnext + 40) < SMEM_TBL;
  const bool quiet = wtot < sp.w1 && !(sp.tic_count >= 0 && sp.tic_count < a.nsamp) && next_fast;
  return !quiet;
}
// parameters of the block after a quiet block, from the closed forms
__device__ __forceinline__ void advance_quiet(StepParams &sp, const TrackArgs &a) {
  const unsigned long long n = (unsigned long long)a.nsamp;
  const unsigned long long kend = (unsigned long long)sp.kph0 + n * sp.kinc;
  const unsigned long long cend = (unsigned long long)sp.cph0 + n * sp.cinc;
  const uint32_t wtot = (uint32_t)(kend >> 32);
  sp.kph0 = (uint32_t)kend;
  sp.cph0 = (uint32_t)cend;
  sp.hc0 += wtot;
  sp.w1 -= wtot;
  sp.cyc_pending += (uint32_t)(cend >> 32);
  if (sp.tic < (long long)n) {
    sp.tic_count = (int)sp.tic;
    sp.tic += a.cfg.tic_ref - (long long)n;
  } else {
    sp.tic_count = -1;
    sp.tic -= (long long)n;
  }
}

// ---- (channel, time-slice) work queue -----------------------------------------------------------------
// A channel's blocks must run in order, but nothing ties a channel to one CTA for the whole record.  The
// launch cuts every channel's nblocks into slices and starts one CTA per (channel, slice) item; a CTA takes
// the next item from a FIFO ticket queue in global memory, runs the slice from the channel state in
// gnssb200_rx (exactly what a second launch would do), stores the state and pushes (channel, slice+1).
// The hardware block scheduler refills an SM as soon as a CTA retires, so all channels advance at the same
// pace and every SM stays full until the end whatever the ratio of channels to SMs (a static
// one-CTA-per-channel grid of 768 CTAs leaves 120 of the 148 SMs at 5 of 6 CTAs, and grids beyond one wave
// leave most of the GPU idle during the last one).  Tickets are handed out in CTA start order and the item
// behind ticket t is pushed by a CTA that holds an earlier ticket, i.e. one that is already running: no
// waiting CTA can depend on one that has not been scheduled.
typedef unsigned long long SchedSlot;  // low word: ticket number of the item stored here, high word: item = channel + nchan * slice
struct SchedQueue {
  unsigned head, tail, total, nchan;
  long long slice_blocks;
  long long *tic;       // [nchan] TIC down-counter of the channel at the start of its next slice
  int32_t *dumpcnt;     // [nchan] dump records written so far (when the caller keeps no counters)
  SchedSlot *slots;     // [nchan]
};

__global__ void sched_init_kernel(SchedQueue *q, unsigned nchan, unsigned nslices, long long slice_blocks, long long *tic,
                                  int32_t *dumpcnt, SchedSlot *slots) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    q->head = 0;
    q->tail = nchan;
    q->total = nchan * nslices;
    q->nchan = nchan;
    q->slice_blocks = slice_blocks;
    q->tic = tic;
    q->dumpcnt = dumpcnt;
    q->slots = slots;
  }
  if (i < nchan) {
    slots[i] = (unsigned long long)i | ((unsigned long long)i << 32);  // ticket i = slice 0 of channel i
    dumpcnt[i] = 0;
    tic[i] = 0;
  }
}

template <class T>
__device__ __forceinline__ void copy_in_cg(T &dst, const T *src) {  // L2-coherent read of state another SM may have written
  static_assert(sizeof(T) % 4 == 0, "word copy");
  const int *s4 = reinterpret_cast<const int *>(src);
  int *d4 = reinterpret_cast<int *>(&dst);
  for (int i = 0; i < (int)(sizeof(T) / 4); i++) d4[i] = __ldcg(s4 + i);
}

// SPT samples per correlator thread: 32 (256 correlator threads, shortest block latency) or 64 (128
// threads: half the per-block overhead instructions and six resident CTAs per SM for dense grids).
template <int MINB, int FMT, int SPT>
__global__ void __launch_bounds__(8192 / SPT + 32, MINB) track_ws_kernel(const TrackArgs a, const int tile_bytes) {
  constexpr int WS_CORR_THREADS = 8192 / SPT;
  constexpr int WS_THREADS = WS_CORR_THREADS + 32;
  constexpr int fmt = FMT;
  constexpr bool packed_native = FMT == GNSSB200_FMT_PACKED2;
  __shared__ ChanShared cs;
  __shared__ BlockParams params[2];
  __shared__ uint2 lut[8];
  __shared__ uint32_t tbl[SMEM_TBL];
  __shared__ uint32_t alias_tbl[2][48];  // per ring slot: tbl[0..47] with entry 0 = the block's stale bits (rule A6)
  __shared__ __align__(16) int totals[12];
  __shared__ __align__(8) uint64_t dfull[2], pfull[2], empty[2], tfull;
  extern __shared__ __align__(128) uint8_t tiles[];
  uint32_t *vlut = reinterpret_cast<uint32_t *>(tiles + 2 * (size_t)tile_bytes);

  __shared__ int s_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t blk_bytes = bytes_for(fmt, a.nsamp);
  constexpr int CTRL = WS_CORR_THREADS;  // the control lane
  SchedQueue *const wq = a.sched;
  // channel-independent tables first: they fill while the control lane may still be waiting for its item
  fill_lo_lut(lut);
  if (packed_native) {
    const int i_lo[8] = {-1, 1, 2, 2, 1, -1, -2, -2};
    const int q_lo[8] = {2, 2, 1, -1, -2, -2, -1, 1};
    const int val[4] = {1, -1, 3, -3};
    for (int i = tid; i < 128 * 32; i += WS_THREADS) {
      const int e = i >> 5, ph = e >> 4, code = e & 15;
      const int I = val[code & 3], Q = val[code >> 2];
      const int ival = i_lo[ph] * I + q_lo[ph] * Q, qval = q_lo[ph] * I - i_lo[ph] * Q;  // correlator.c:214-215
      vlut[i] = (uint32_t)(ival + 65536 * qval);
    }
  }
  // One item per channel (few channels, or a short run): no queue at all, the CTA index is the channel.
  const bool queued = wq != nullptr;
  if (queued && tid == CTRL) {  // this CTA's item
    const unsigned ticket = atomicAdd(&wq->head, 1u);
    volatile SchedSlot *slot = wq->slots + ticket % wq->nchan;
    unsigned long long v;
    while ((unsigned)(v = *slot) != ticket) __nanosleep(100);
    __threadfence();  // acquire: the state the previous slice of this channel stored
    s_item = (int)(v >> 32);
  }
  __syncthreads();
  const int item = queued ? s_item : (int)blockIdx.x;
  const int nchan = queued ? (int)wq->nchan : (int)gridDim.x;
  const long long slice_blocks = queued ? wq->slice_blocks : a.nblocks;
  const int chan_id = item % nchan, slice = item / nchan;
  const int s = a.first_stream + chan_id / NCH, ch = chan_id % NCH;
  gnssb200_rx *rx = a.rx + s;
  const int tbl_prn = __ldcg(&rx->reg_write[ch << 3]);
  const long long slice_first = (long long)slice * slice_blocks;          // first block of this slice within the launch
  const long long nblocks = min(slice_blocks, a.nblocks - slice_first);    // blocks of this slice
  const uint8_t *stream_base = a.d_if + (size_t)s * a.stride + (size_t)slice_first * blk_bytes;

  if (tid < 12) totals[tid] = 0;
  for (int i = tid; i < SMEM_TBL; i += WS_THREADS) {
    long long f = (long long)tbl_prn * HALF_CHIPS + i;
    tbl[i] = (tbl_prn >= 1 && tbl_prn <= 32 && f < TABLE_ENTRIES) ? a.code_table[f] : 0u;
  }
  if (tid < 96) {
    const int t = tid % 48;
    long long f = (long long)tbl_prn * HALF_CHIPS + t;
    alias_tbl[tid / 48][t] = (tbl_prn >= 1 && tbl_prn <= 32 && f < TABLE_ENTRIES) ? a.code_table[f] : 0u;
  }
  __syncthreads();  // tbl complete before the control lane looks up stale bits
  const int b8 = ch << 3;
  StepParams sp;
  long long first_block = 0;
  long long loaded = -1;  // last block whose TMA load was issued
  if (tid == CTRL) {
    copy_in_cg(cs.k, &rx->chan[ch]);
    copy_in_cg(cs.g, &rx->corr[ch]);
    cs.r.w_prn = __ldcg(&rx->reg_write[b8]);
    cs.r.w_carr_hi = __ldcg(&rx->reg_write[b8 + 3]);
    cs.r.w_carr_lo = __ldcg(&rx->reg_write[b8 + 4]);
    cs.r.w_code_hi = __ldcg(&rx->reg_write[b8 + 5]);
    cs.r.w_code_lo = __ldcg(&rx->reg_write[b8 + 6]);
    cs.r.w_epoch = __ldcg(&rx->reg_write[b8 + 7]);
    cs.r.w_slew = __ldcg(&rx->reg_write[b8 + 0x84]);
    for (int j = 0; j < 8; j++) cs.r.r_meas[j] = __ldcg(&rx->reg_read[b8 + j]);
    for (int j = 0; j < 6; j++) cs.r.r_acc[j] = __ldcg(&rx->reg_read[b8 + 0x84 + j]);
    const int prev_flags = slice > 0 ? __ldcg(&a.chan_flags[s * NCH + ch]) : 0;
    cs.tic = slice > 0 ? __ldcg(&wq->tic[chan_id]) : rx->tic;
    cs.dumped_last = prev_flags & 1;
    cs.halted = (prev_flags >> 1) & 1;
    cs.dump_count = a.dump_count ? __ldcg(&a.dump_count[s * NCH + ch]) : (slice > 0 ? __ldcg(&wq->dumpcnt[chan_id]) : 0);
    first_block = rx->blocks_done + slice_first;
    sp.stale_bits = 0;
    if (nblocks > 0 && !rx->halted && !cs.halted) {
      prepare_block(cs, sp, a, tbl_prn);
      if (sp.mode == MODE_FAST) sp.stale_bits = tbl[sp.stale_idx];
    } else
      sp.mode = MODE_STOP;
    mbar_init(&dfull[0], 1);
    mbar_init(&dfull[1], 1);
    mbar_init(&pfull[0], 1);
    mbar_init(&pfull[1], 1);
    mbar_init(&empty[0], WS_CORR_THREADS / 32);
    mbar_init(&empty[1], WS_CORR_THREADS / 32);
    mbar_init(&tfull, WS_CORR_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (sp.mode == MODE_FAST || sp.mode == MODE_SERIAL) {
      mbar_expect_tx(&dfull[0], (uint32_t)blk_bytes);
      tma_load_1d(tiles, stream_base, (uint32_t)blk_bytes, &dfull[0]);
      loaded = 0;
    }
  }
  __syncthreads();  // mbarriers initialised

  // ---------------- control lane ----------------
  if (warp == WS_CORR_THREADS / 32) {
    if (lane != 0) return;
    auto publish = [&](int slot, const StepParams &sp, bool event) {
      BlockParams &p = params[slot];
      p.cph0 = sp.cph0; p.kph0 = sp.kph0; p.cinc = sp.cinc; p.kinc = sp.kinc;
      p.hc0 = sp.hc0; p.w1 = sp.w1; p.stale_idx = sp.stale_idx; p.stale_bits = sp.stale_bits;
      p.mode = sp.mode; p.event = event ? 1 : 0;
      alias_tbl[slot][0] = sp.stale_bits;
      mbar_arrive(&pfull[slot]);  // release: the stores above are visible to whoever observes the phase
    };
    bool event = block_is_event(sp, a, nblocks <= 1);
    publish(0, sp, event);
    uint32_t ev_phase = 0;
#ifdef TRACK_PROFILE
    long long c_twait = 0, c_fin = 0, c_words = 0, c_params = 0, c_rest = 0, c_ewait = 0, c_quiet = 0, n_ev = 0, n_q = 0, c_acc = 0, c_prep = 0, c_cls = 0;
#define CP(var) { long long _c = clock64(); var += _c - _t; _t = _c; }
#else
#define CP(var)
#endif
    for (long long b = 0; b < nblocks; b++) {
      if (sp.mode == MODE_STOP) break;
#ifdef TRACK_PROFILE
      long long _t = clock64();
#endif
      const bool last = b + 1 == nblocks;
      const int slot = (int)(b & 1), nslot = slot ^ 1;
      if (!last) {
        // ring slot of block b+1 (parameters, alias table, tile) is free once every warp finished block b-1
        if (b >= 1) mbar_wait(&empty[nslot], (uint32_t)(((b - 1) >> 1) & 1));
        if (sp.mode != MODE_IDLE) {
          mbar_expect_tx(&dfull[nslot], (uint32_t)blk_bytes);
          tma_load_1d(tiles + (size_t)nslot * tile_bytes, stream_base + (size_t)(b + 1) * blk_bytes, (uint32_t)blk_bytes, &dfull[nslot]);
          loaded = b + 1;
        }
      }
      CP(c_ewait)
      if (!event) {  // quiet block: nothing leaves the correlator threads
        advance_quiet(sp, a);
        event = block_is_event(sp, a, b + 2 == nblocks);
        publish(nslot, sp, event);
#ifdef TRACK_PROFILE
        n_q++;
#endif
        CP(c_quiet)
        continue;
      }
#ifdef TRACK_PROFILE
      n_ev++;
#endif
      // what follows from the block's parameters alone is settled while the correlator warps still work
      cs.tic = sp.tic;
      cs.g.carrier_cycle += sp.cyc_pending;
      const int was_mode = sp.mode;
      if (was_mode == MODE_FAST) {
        finalize_state(cs, sp, a.nsamp);
        if (!last) prepare_block_state(cs, sp, a);  // sp now describes block b+1 as far as the correlator state decides it
      }
      CP(c_fin)
      if (was_mode == MODE_FAST) {
        int A[6], B[6];
        mbar_wait(&tfull, ev_phase);
        ev_phase ^= 1;
        CP(c_twait)
        const int4 t0 = *reinterpret_cast<const int4 *>(&totals[0]);
        const int4 t1 = *reinterpret_cast<const int4 *>(&totals[4]);
        const int4 t2 = *reinterpret_cast<const int4 *>(&totals[8]);
        A[0] = t0.x; A[1] = t0.y; A[2] = t0.z; A[3] = t0.w; A[4] = t1.x; A[5] = t1.y;
        B[0] = t1.z; B[1] = t1.w; B[2] = t2.x; B[3] = t2.y; B[4] = t2.z; B[5] = t2.w;
        const int4 z = make_int4(0, 0, 0, 0);
        *reinterpret_cast<int4 *>(&totals[0]) = z;  // the next event block's atomics come after the publish below
        *reinterpret_cast<int4 *>(&totals[4]) = z;
        *reinterpret_cast<int4 *>(&totals[8]) = z;
        finalize_acc(cs, A, B);
        CP(c_acc)
      } else if (was_mode == MODE_SERIAL) {
        mbar_wait(&dfull[slot], (uint32_t)((b >> 1) & 1));
        serial_block(cs, sp, a.code_table, fmt, a.nsamp, tiles + (size_t)slot * tile_bytes);
        if (!last) prepare_block_state(cs, sp, a);
      } else
        cs.dumped_last = 0;
      // ISR, first part: whatever can change the NCO words / slew
      int st_in = -1;
      bool isr = false;
      if (cs.dumped_last && a.run_isr) {
        if (dev_gpsisr_words(cs.k, cs.r, a.cfg, st_in))
          cs.halted = 1;
        else
          isr = true;
      }
      CP(c_words)
      if (cs.halted || was_mode == MODE_IDLE)  // an idle channel has no ISR: nothing can change any more
        sp.mode = MODE_STOP;
      else if (!last) {
        prepare_block_regs(cs, sp, a, tbl_prn);
        sp.stale_bits = sp.mode == MODE_FAST ? tbl[sp.stale_idx] : 0u;
      }
      CP(c_prep)
      if (!last) {
        event = block_is_event(sp, a, b + 2 == nblocks);
        CP(c_cls)
        publish(nslot, sp, event);
      }
      CP(c_params)
      // second part, off the correlators' critical path
      if (isr) dev_gpsisr_rest(cs.k, cs.r, a.cfg, st_in);
      if (cs.dumped_last && !cs.halted && a.dumps && cs.dump_count < a.dump_cap) {
        gnssb200_dump *out = &a.dumps[((size_t)s * NCH + ch) * a.dump_cap + cs.dump_count];
        int4 q0, q1, q2;
        q0.x = (int)(first_block + b);
        q0.y = (int)(uint16_t)(int16_t)ch | ((int)(uint16_t)(int16_t)cs.k.state << 16);
        q0.z = cs.r.r_acc[0];
        q0.w = cs.r.r_acc[1];
        q1.x = cs.r.r_acc[2];
        q1.y = cs.r.r_acc[3];
        q1.z = cs.r.r_acc[4];
        q1.w = cs.r.r_acc[5];
        q2.x = (cs.r.w_carr_hi << 16) + cs.r.w_carr_lo;
        q2.y = (cs.r.w_code_hi << 16) + cs.r.w_code_lo;
        q2.z = (int)(uint16_t)(int16_t)cs.k.n_freq | ((int)(uint16_t)(int16_t)cs.k.codes << 16);
        q2.w = cs.r.w_slew;
        int4 *o4 = reinterpret_cast<int4 *>(out);
        o4[0] = q0;
        o4[1] = q1;
        o4[2] = q2;
        cs.dump_count++;
      }
      if (!last && sp.mode != MODE_STOP) apply_epoch_load(cs);  // start-of-block rule of the next block
      CP(c_rest)
    }
#ifdef TRACK_PROFILE
    if (blockIdx.x == 0 && n_ev && n_q)
      printf("control lane: %lld quiet blocks: slot wait+TMA %lld, classify+publish %lld | %lld event blocks: totals wait %lld finalize %lld isr words %lld params+publish %lld rest %lld (cycles each)\n",
             n_q, c_ewait / (n_q + n_ev), c_quiet / n_q, n_ev, c_twait / n_ev, c_fin / n_ev, c_words / n_ev, c_params / n_ev, c_rest / n_ev);
    if (blockIdx.x == 0 && n_ev)
      printf("   after the totals: read+accumulators %lld, isr words %lld, prepare params %lld, classify %lld, publish %lld\n", c_acc / n_ev, c_words / n_ev, c_prep / n_ev,
             c_cls / n_ev, c_params / n_ev);
#endif
    // a prefetched block nobody consumed must land before the CTA may exit
    if (loaded >= 0) mbar_wait(&dfull[loaded & 1], (uint32_t)((loaded >> 1) & 1));
    rx->chan[ch] = cs.k;
    rx->corr[ch] = cs.g;
    rx->reg_write[b8 + 3] = cs.r.w_carr_hi;
    rx->reg_write[b8 + 4] = cs.r.w_carr_lo;
    rx->reg_write[b8 + 5] = cs.r.w_code_hi;
    rx->reg_write[b8 + 6] = cs.r.w_code_lo;
    rx->reg_write[b8 + 7] = cs.r.w_epoch;
    rx->reg_write[b8 + 0x84] = cs.r.w_slew;
    for (int q = 1; q < 8; q++) rx->reg_read[b8 + q] = cs.r.r_meas[q];
    for (int q = 0; q < 6; q++) rx->reg_read[b8 + 0x84 + q] = cs.r.r_acc[q];
    a.chan_flags[s * NCH + ch] = (cs.dumped_last ? 1 : 0) | (cs.halted ? 2 : 0);
    if (a.dump_count) a.dump_count[s * NCH + ch] = cs.dump_count;
    if (queued) {
      if (!a.dump_count) wq->dumpcnt[chan_id] = cs.dump_count;
      wq->tic[chan_id] = cs.tic;
    }
    const unsigned next_item = (unsigned)item + (unsigned)nchan;  // the channel's next slice
    if (queued && next_item < wq->total) {
      __threadfence();  // release: the state stored above, before the item becomes visible
      const unsigned t = atomicAdd(&wq->tail, 1u);
      atomicExch(wq->slots + t % wq->nchan, (unsigned long long)t | ((unsigned long long)next_item << 32));
    }
    return;
  }

  // ---------------- correlator warps ----------------
  const int i0 = tid * SPT;
  const bool live = i0 < a.nsamp;
  const uint32_t vlut_lane = smem_u32(vlut) + 4u * (uint32_t)lane;
  int carry[6] = {0, 0, 0, 0, 0, 0};
#ifdef TRACK_PROFILE
  long long t_pw = 0, t_dw = 0, t_corr = 0, t_red = 0, t_all = -clock64(), nb = 0;
#endif
  for (long long b = 0; b < nblocks; b++) {
    const int slot = (int)(b & 1);
    const uint32_t par = (uint32_t)((b >> 1) & 1);
#ifdef TRACK_PROFILE
    long long _t = clock64();
    nb++;
#endif
    mbar_wait(&pfull[slot], par);
    CP(t_pw)
    const uint4 p0 = reinterpret_cast<const uint4 *>(&params[slot])[0];
    const uint4 p1 = reinterpret_cast<const uint4 *>(&params[slot])[1];
    const int2 p2 = reinterpret_cast<const int2 *>(&params[slot])[4];
    const int mode = p2.x;
    const bool event = p2.y != 0;
    if (mode == MODE_STOP) break;
    if (mode == MODE_FAST) {
      const uint32_t cph0 = p0.x, kph0 = p0.y, cinc = p0.z, kinc = p0.w;
      const uint32_t hc0 = p1.x, w1 = p1.y, stale_idx = p1.z, stale_bits = p1.w;
      const uint8_t *tile = tiles + (size_t)slot * tile_bytes;
      mbar_wait(&dfull[slot], par);
      CP(t_dw)
      int sumA[6] = {0, 0, 0, 0, 0, 0}, sumB[6] = {0, 0, 0, 0, 0, 0};
      bool anyB = false;
      {
        uint32_t cur[SPT / 2];
        uint32_t pk[SPT / 8];
        if (live && packed_native) {
          const uint32_t *pp = reinterpret_cast<const uint32_t *>(tile + (i0 >> 1));
#pragma unroll
          for (int q = 0; q < SPT / 8; q++) pk[q] = pp[q];
        } else if (live) {
          load_chunk<SPT, true>(tile, fmt, i0, a.nsamp, true, nullptr, cur);
        }
        const int i1 = live ? min(i0 + SPT, a.nsamp) : i0 + 1;
        const unsigned long long k0 = (unsigned long long)kph0 + (unsigned long long)i0 * kinc;
        const uint32_t w_start = (uint32_t)(k0 >> 32);
        const uint32_t w_lastb = (uint32_t)(((unsigned long long)kph0 + (unsigned long long)(i1 - 1) * kinc) >> 32);
        const bool allA = !live || w_lastb < w1, allB = live && w_start >= w1;
        uint32_t h, hl;
        if (allB) {
          h = w_start - w1;
          hl = (h == 0) ? stale_idx : h;  // stale bits after the dump (SURVEY.md App. A rule A6)
        } else {
          h = hc0 + w_start;
          hl = h;
        }
        int pE = 0, pP = 0, pL = 0;
        // chunk starting in the first post-dump half chip: stale bits first, then tbl[1], tbl[2], ...
        const bool stale_start = allB && h == 0;
        if (live && packed_native)
          correlate_chunk_packed<SPT>(pk, cph0 + (uint32_t)i0 * cinc, (uint32_t)k0, cinc, kinc, stale_start ? alias_tbl[slot] : tbl, h,
                                      stale_start ? stale_bits : tbl[hl], vlut_lane, PipeK{a.k1, a.k8, a.k128, a.k2048}, pE, pP, pL);
        else if (live)
          correlate_chunk<SPT>(cur, cph0 + (uint32_t)i0 * cinc, (uint32_t)k0, cinc, kinc, stale_start ? alias_tbl[slot] : tbl, h,
                               stale_start ? stale_bits : tbl[hl], lut, pE, pP, pL);
        const bool straddle = !allA && !allB;
        if (!straddle && live) {
          int v[6];
          unpack_lanes(pL, v[0], v[1]);
          unpack_lanes(pP, v[2], v[3]);
          unpack_lanes(pE, v[4], v[5]);
          if (allA) {
#pragma unroll
            for (int q = 0; q < 6; q++) sumA[q] += v[q];
          } else {
#pragma unroll
            for (int q = 0; q < 6; q++) sumB[q] += v[q];
          }
        }
        // the chunk that contains the dump is re-evaluated one sample per lane by its warp
        unsigned m = __ballot_sync(0xffffffffu, straddle);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const int si0 = __shfl_sync(0xffffffffu, i0, src);
          for (int i = si0 + lane; i < min(si0 + SPT, a.nsamp); i += 32) {
            const unsigned long long ki = (unsigned long long)kph0 + (unsigned long long)i * kinc;
            const uint32_t wb = (uint32_t)(ki >> 32);
            const bool inA = wb < w1;
            const uint32_t rel = wb - w1;
            const uint32_t hh = inA ? hc0 + wb : (rel == 0 ? stale_idx : rel);
            const uint32_t t = tbl[hh];
            int I, Q;
            load_sample(tile, fmt, i, I, Q);
            const uint2 ab = lut[(cph0 + (uint32_t)i * cinc) >> 29];
            const int v = I * (int)ab.x + Q * (int)ab.y;
            int vi, vq;
            unpack_lanes(v, vi, vq);
            const int cE = sext8(t, 0), cP = sext8(t, 1), cL = sext8(t, 2);
            if (inA) {
              sumA[0] += cL * vi; sumA[1] += cL * vq; sumA[2] += cP * vi;
              sumA[3] += cP * vq; sumA[4] += cE * vi; sumA[5] += cE * vq;
            } else {
              sumB[0] += cL * vi; sumB[1] += cL * vq; sumB[2] += cP * vi;
              sumB[3] += cP * vq; sumB[4] += cE * vi; sumB[5] += cE * vq;
            }
          }
        }
        anyB |= !allA;
      }
      CP(t_corr)
      if (!event) {  // no dump in this block: every chunk was pre-dump, keep the sums in registers
#pragma unroll
        for (int q = 0; q < 6; q++) carry[q] += sumA[q];
      } else {
        const bool warpB = __any_sync(0xffffffffu, anyB);
        int va = 0, vb = 0;
#pragma unroll
        for (int q = 0; q < 6; q++) {
          const int ra = warp_sum(sumA[q] + carry[q]);
          carry[q] = 0;
          if (lane == q) va = ra;
        }
        if (warpB) {
#pragma unroll
          for (int q = 0; q < 6; q++) {
            const int rb = warp_sum(sumB[q]);
            if (lane == q) vb = rb;
          }
        }
        if (lane < 6) {
          atomicAdd(&totals[lane], va);
          if (warpB) atomicAdd(&totals[6 + lane], vb);
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      if (event && mode == MODE_FAST) mbar_arrive(&tfull);
      mbar_arrive(&empty[slot]);
    }
    CP(t_red)
  }
#ifdef TRACK_PROFILE
  t_all += clock64();
  if (blockIdx.x == 0 && (tid == 0 || tid == 133) && nb)
    printf("correlator tid %d: per block: params wait %lld  data wait %lld  load+correlate+post %lld  reduce/arrive %lld  total %lld\n", tid, t_pw / nb, t_dw / nb,
           t_corr / nb, t_red / nb, t_all / nb);
#endif
}

// one thread per stream: status words, TIC counter and block counter after a run
__global__ void track_finish_kernel(gnssb200_rx *rx, const int32_t *chan_flags, int first_stream, int n_streams,
                                    int nsamp, long long nblocks, long long tic_ref) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_streams) return;
  const int s = first_stream + t;
  gnssb200_rx *r = rx + s;
  if (nblocks <= 0 || r->halted) return;
  int status = 0, halted = 0;
  for (int ch = 0; ch < NCH; ch++) {
    const int f = chan_flags[s * NCH + ch];
    if (f & 1) status |= 1 << ch;
    if (f & 2) halted = 1;
  }
  long long tic = r->tic;
  int tic_count = -1;
  for (long long b = 0; b < nblocks; b++) {
    if (tic < nsamp) {
      tic_count = (int)tic;
      tic += tic_ref - nsamp;
    } else {
      tic -= nsamp;
      tic_count = -1;
    }
  }
  r->tic = tic;
  r->reg_read[0x82] = status;                        // correlator.c:309
  r->reg_read[0x83] = (tic_count > -1) ? 0x2000 : 0; // :312-315
  r->blocks_done += nblocks;
  if (halted) r->halted = 1;
}

// ---- diagnostics: the device ISR's integer helpers on arrays (tests/test_isr_math.py) -----------------
__global__ void isr_math_kernel(int n, const int *y, const int *x, int *at, const long long *L, unsigned *sq, const int *num,
                                const int *den, int *dv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  at[i] = dev_atan2_i32(y[i], x[i]);
  sq[i] = dev_isqrt(L[i]);
  dv[i] = dev_div_small(num[i], den[i]);
}

extern "C" int gnssb200_isr_math_eval(gnssb200_handle *h, int n, const int32_t *y, const int32_t *x, int32_t *atan_out,
                                      const int64_t *L, uint32_t *sqrt_out, const int32_t *num, const int32_t *den,
                                      int32_t *div_out) {
  if (!h || n <= 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  int *d_i = nullptr;
  long long *d_L = nullptr;
  CUDA_TRY(cudaMalloc(&d_i, (size_t)n * 7 * sizeof(int)));
  CUDA_TRY(cudaMalloc(&d_L, (size_t)n * sizeof(long long)));
  int *dy = d_i, *dx = d_i + n, *dat = d_i + 2 * (size_t)n, *dnum = d_i + 3 * (size_t)n, *dden = d_i + 4 * (size_t)n,
      *ddv = d_i + 5 * (size_t)n;
  unsigned *dsq = (unsigned *)(d_i + 6 * (size_t)n);
  CUDA_TRY(cudaMemcpy(dy, y, (size_t)n * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dx, x, (size_t)n * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dnum, num, (size_t)n * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dden, den, (size_t)n * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d_L, L, (size_t)n * 8, cudaMemcpyHostToDevice));
  isr_math_kernel<<<(n + 255) / 256, 256>>>(n, dy, dx, dat, d_L, dsq, dnum, dden, ddv);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(atan_out, dat, (size_t)n * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(sqrt_out, dsq, (size_t)n * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(div_out, ddv, (size_t)n * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_i);
  cudaFree(d_L);
  h->launches += 1;
  return 0;
}

// ---- host side -----------------------------------------------------------------------------------
void build_code_table_host(uint32_t *table) {
  // C/A Gold codes, G2 register start states per PRN (IS-GPS-200; same values as correlator.c:67-71),
  // replicas at half-chip spacing: early[h]=c[h>>1], prompt[h]=c[((h+1)%2046)>>1], late[h]=c[((h+2)%2046)>>1]
  static const int g2_start[33] = {0x000, 0x3f6, 0x3ec, 0x3d8, 0x3b0, 0x04b, 0x096, 0x2cb, 0x196, 0x32c, 0x3ba,
                                   0x374, 0x1d0, 0x3a0, 0x340, 0x280, 0x100, 0x113, 0x226, 0x04c, 0x098, 0x130,
                                   0x260, 0x267, 0x338, 0x270, 0x0e0, 0x1c0, 0x380, 0x22b, 0x056, 0x0ac, 0x158};
  for (int i = 0; i <= TABLE_ENTRIES; i++) table[i] = 0;
  for (int prn = 1; prn <= 32; prn++) {
    int chip[1023];
    int g1 = 0x1FF, g2 = g2_start[prn];
    chip[0] = 1;
    for (int c = 1; c < 1023; c++) {
      chip[c] = (g1 ^ g2) & 1;
      g1 = (g1 >> 1) | (((g1 << 2) ^ (g1 << 9)) & 0x200);
      g2 = (g2 >> 1) | (((g2 << 1) ^ (g2 << 2) ^ (g2 << 5) ^ (g2 << 7) ^ (g2 << 8) ^ (g2 << 9)) & 0x200);
    }
    for (int h = 0; h < HALF_CHIPS; h++) {
      const int e = 2 * chip[(h % HALF_CHIPS) >> 1] - 1;
      const int p = 2 * chip[((h + 1) % HALF_CHIPS) >> 1] - 1;
      const int l = 2 * chip[((h + 2) % HALF_CHIPS) >> 1] - 1;
      table[prn * HALF_CHIPS + h] = (uint32_t)(e & 0xff) | ((uint32_t)(p & 0xff) << 8) | ((uint32_t)(l & 0xff) << 16);
    }
  }
}

size_t track_sched_bytes(int n_streams) {  // work-queue storage for a handle with n_streams receivers
  return sizeof(SchedQueue) * (size_t)n_streams + (size_t)n_streams * NCH * (8 + 8 + 4) + 256;
}

int track_launch(gnssb200_handle *h, int first_stream, int n_streams, const void *d_if, size_t stride, int fmt,
                 int nsamp, long long nblocks, int run_isr, gnssb200_dump *d_dumps, int dump_cap,
                 int32_t *d_dump_count, cudaStream_t st) {
  if (n_streams <= 0 || nblocks <= 0) return 0;
  TrackArgs a;
  a.rx = h->d_rx;
  a.chan_flags = h->d_chan_flags;
  a.code_table = h->d_code_table;
  a.d_if = (const uint8_t *)d_if - (size_t)first_stream * stride;  // kernel indexes by absolute stream
  a.stride = stride;
  a.fmt = fmt;
  a.nsamp = nsamp;
  a.nblocks = nblocks;
  a.run_isr = run_isr;
  a.first_stream = first_stream;
  a.dumps = d_dumps;
  a.dump_cap = dump_cap;
  a.dump_count = d_dump_count;
  a.sched = nullptr;
  a.k1 = 1u;
  a.k8 = 8u;
  a.k128 = 128u;
  a.k2048 = 2048u;
  static_cast<gnssb200_cfg &>(a.cfg) = h->cfg;
  {
    const double m = h->cfg.clock_mult;
    const long long im = (long long)m;
    a.cfg.mult_i = ((double)im == m && im > -1024 && im < 1024) ? (int)im : 0;
    const gnssb200_cfg &c = h->cfg;
    auto ab = [](long long v) { return v < 0 ? -v : v; };
    const bool pll_ok = (ab(c.pll_i1) + ab(c.pll_i2)) * (1ll << 17) + ab(c.pll_i3) * (1ll << 16) < (1ll << 31);
    const bool dll_ok = (ab((long long)c.dll_i1 + 1) + ab(c.dll_i2)) * (1ll << 17) < (1ll << 31);
    const int shc = 32 - c.carrier_nco_bits, shk = 32 - c.code_nco_bits;
    a.cfg.fast32 = (a.cfg.mult_i != 0 && pll_ok && dll_ok && shc >= 0 && shc <= 8 && shk >= 0 && shk <= 8) ? 1 : 0;
  }
  const int grid = n_streams * NCH;
  const int spt = 32;
  int threads = (nsamp + spt - 1) / spt;  // track_loop_kernel: 32 samples per thread per pass
  threads = ((threads + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  // TMA staging needs 16-byte aligned blocks that fit one tile per CTA pass
  const size_t blk_bytes = fmt == GNSSB200_FMT_INT8_IQ ? (size_t)nsamp * 2 : (fmt == GNSSB200_FMT_PACKED2 ? (size_t)nsamp / 2 : (size_t)nsamp);
  const bool aligned = (((uintptr_t)d_if | stride | blk_bytes) & 15) == 0 && (nsamp % 8) == 0;
  static int no_tma = -1;
  if (no_tma < 0) {
    const char *e = getenv("GNSSB200_TRACK_NO_TMA");
    no_tma = (e && atoi(e)) ? 1 : 0;
  }
  const int use_tma = (aligned && !no_tma && nsamp <= 256 * spt && blk_bytes <= 16384) ? 1 : 0;
  const int tile_bytes = use_tma ? (int)((blk_bytes + 127) & ~(size_t)127) : 0;
  const size_t dyn = (size_t)2 * tile_bytes + ((use_tma && fmt == GNSSB200_FMT_PACKED2) ? 128 * 32 * 4 : 0);
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(track_loop_kernel<256, 2, GNSSB200_FMT_INT8_IQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384 + 256));
    CUDA_TRY(cudaFuncSetAttribute(track_loop_kernel<256, 2, GNSSB200_FMT_PACKED2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16384 + 256));
    attr_done = true;
  }
  static int force_occ = -1;  // GNSSB200_TRACK_OCC: force the resident-CTAs-per-SM variant (experiments)
  if (force_occ < 0) {
    const char *e = getenv("GNSSB200_TRACK_OCC");
    force_occ = e ? atoi(e) : 0;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const bool hot = use_tma && nsamp <= 8192 && (fmt == GNSSB200_FMT_INT8_IQ || fmt == GNSSB200_FMT_PACKED2);
  static int use_ws = -1;
  if (use_ws < 0) {
    const char *e = getenv("GNSSB200_TRACK_WS");
    use_ws = e ? atoi(e) : 1;
  }
  if (use_ws && hot) {  // warp-specialised variant: correlator warps + control lane
    constexpr int DSM_I8 = 2 * 16384 + 256, DSM_PK = 3 * 16384 + 256;
    static bool aws = false;
    if (!aws) {
      CUDA_TRY(cudaFuncSetAttribute(track_ws_kernel<2, GNSSB200_FMT_INT8_IQ, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, DSM_I8));
      CUDA_TRY(cudaFuncSetAttribute(track_ws_kernel<3, GNSSB200_FMT_INT8_IQ, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, DSM_I8));
      CUDA_TRY(cudaFuncSetAttribute(track_ws_kernel<2, GNSSB200_FMT_PACKED2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, DSM_PK));
      CUDA_TRY(cudaFuncSetAttribute(track_ws_kernel<3, GNSSB200_FMT_PACKED2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, DSM_PK));
      CUDA_TRY(cudaFuncSetAttribute(track_ws_kernel<4, GNSSB200_FMT_PACKED2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DSM_PK));
      CUDA_TRY(cudaFuncSetAttribute(track_ws_kernel<6, GNSSB200_FMT_PACKED2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DSM_PK));
      CUDA_TRY(cudaFuncSetAttribute(track_ws_kernel<5, GNSSB200_FMT_PACKED2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DSM_PK));
      aws = true;
    }
    // Work queue: every channel's blocks are cut into slices of slice_blocks; one CTA per (channel, slice)
    // item, items handed out through the FIFO so that a channel's slices run in order.
    static long long env_slice = -1;
    if (env_slice < 0) {
      const char *e = getenv("GNSSB200_TRACK_SLICE");
      env_slice = e ? atoll(e) : 0;
    }
    // Slicing pays where the GPU is throughput bound (four or more channels per SM): there it keeps every SM
    // full to the end.  With few channels each channel's latency is the limit, tickets land on SMs at random
    // (two running CTAs may share an SM next to an idle one), so those runs stay one item per channel.
    const int per_sm_need = (grid + sms - 1) / sms;
    const long long slice_blocks = h->track_slice > 0 ? h->track_slice : (env_slice > 0 ? env_slice : (per_sm_need >= 4 ? 128 : nblocks));
    const unsigned nslices = (unsigned)((nblocks + slice_blocks - 1) / slice_blocks);
    const size_t n_all = (size_t)h->n_streams * NCH;
    uint8_t *base = (uint8_t *)h->d_sched;
    SchedQueue *qd = reinterpret_cast<SchedQueue *>(base) + first_stream;  // one header per possible first stream
    long long *tic = reinterpret_cast<long long *>(base + sizeof(SchedQueue) * (size_t)h->n_streams) + (size_t)first_stream * NCH;
    SchedSlot *slots = reinterpret_cast<SchedSlot *>(base + sizeof(SchedQueue) * (size_t)h->n_streams + 8 * n_all) + (size_t)first_stream * NCH;
    int32_t *dcnt = reinterpret_cast<int32_t *>(base + sizeof(SchedQueue) * (size_t)h->n_streams + 16 * n_all) + (size_t)first_stream * NCH;
    if (nslices > 1) {
      sched_init_kernel<<<(grid + 255) / 256, 256, 0, st>>>(qd, (unsigned)grid, nslices, slice_blocks, tic, dcnt, slots);
      CUDA_TRY(cudaGetLastError());
      h->launches += 1;
      a.sched = qd;
    }
    const unsigned items = (unsigned)grid * nslices;
    // CTAs per SM the channels ask for -> variant (registers / samples per thread)
    const int per_sm = force_occ ? force_occ : (grid + sms - 1) / sms;
    if (fmt == GNSSB200_FMT_INT8_IQ && per_sm >= 3)
      track_ws_kernel<3, GNSSB200_FMT_INT8_IQ, 32><<<items, 288, dyn, st>>>(a, tile_bytes);
    else if (fmt == GNSSB200_FMT_INT8_IQ)
      track_ws_kernel<2, GNSSB200_FMT_INT8_IQ, 32><<<items, 288, dyn, st>>>(a, tile_bytes);
    else if (force_occ == 6)
      track_ws_kernel<6, GNSSB200_FMT_PACKED2, 64><<<items, 160, dyn, st>>>(a, tile_bytes);
    else if (per_sm >= 5)  // five resident CTAs of 72 registers beat six of 64 (spills) by 1-4 % under the work queue
      track_ws_kernel<5, GNSSB200_FMT_PACKED2, 64><<<items, 160, dyn, st>>>(a, tile_bytes);
    else if (per_sm == 4)
      track_ws_kernel<4, GNSSB200_FMT_PACKED2, 64><<<items, 160, dyn, st>>>(a, tile_bytes);
    else if (per_sm == 3)
      track_ws_kernel<3, GNSSB200_FMT_PACKED2, 32><<<items, 288, dyn, st>>>(a, tile_bytes);
    else
      track_ws_kernel<2, GNSSB200_FMT_PACKED2, 32><<<items, 288, dyn, st>>>(a, tile_bytes);
  } else
  if (hot && fmt == GNSSB200_FMT_INT8_IQ)  // GNSSB200_TRACK_WS=0: the barrier-synchronised predecessor, kept for A/B runs
    track_loop_kernel<256, 2, GNSSB200_FMT_INT8_IQ, true><<<grid, threads, dyn, st>>>(a, tile_bytes);
  else if (hot)
    track_loop_kernel<256, 2, GNSSB200_FMT_PACKED2, true><<<grid, threads, dyn, st>>>(a, tile_bytes);
  else
    track_loop_kernel<1024, 1, -1, false><<<grid, threads, 0, st>>>(a, 0);
  CUDA_TRY(cudaGetLastError());
  track_finish_kernel<<<(n_streams + 127) / 128, 128, 0, st>>>(h->d_rx, h->d_chan_flags, first_stream, n_streams, nsamp,
                                                               nblocks, h->cfg.tic_ref);
  CUDA_TRY(cudaGetLastError());
  h->launches += 2;
  return 0;
}
