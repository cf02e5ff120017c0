// track_ws.cuh -- track_ws_kernel: the warp-specialised channel loop (the hot path) and its (channel, slice) work queue.
#pragma once
#include "track_common.cuh"
#include "track_seg.cuh"

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

// Correlator side, two forms:
//   SEGH == 0  8192/CT consecutive samples per correlator thread: 32 (CT = 256, shortest block latency) or 64 (CT =
//              128: half the per-block overhead instructions, more resident CTAs per SM).  Any code NCO rate
//              with 4*kinc < 2^32, both sample formats.
//   SEGH  > 0  SEGH consecutive half-chip segments per correlator thread (track_seg.cuh), CT correlator
//              threads: packed input, code NCO rates with 7 or 8 samples per half chip (the GPS C/A code at the
//              front end's 16 Msps); blocks with another rate are evaluated sample by sample from the closed forms.
template <int MINB, int FMT, int CT, int SEGH = 0>
__global__ void __launch_bounds__(CT + 32, MINB) track_ws_kernel(const TrackArgs a, const int tile_bytes) {
  constexpr int WS_CORR_THREADS = CT;
  constexpr int SPT = SEGH > 0 ? 32 : 8192 / CT;  // samples per correlator thread of the fixed-run form
  static_assert(SEGH > 0 || CT * SPT == 8192, "fixed-run form: CT threads x SPT samples cover the largest block");
  constexpr int WS_THREADS = WS_CORR_THREADS + 32;
  constexpr int fmt = FMT;
  constexpr bool packed_native = FMT == GNSSB200_FMT_PACKED2;
  static_assert(SEGH == 0 || FMT == GNSSB200_FMT_PACKED2 || FMT == GNSSB200_FMT_INT8_IQ, "the segment form reads packed or int8 I,Q samples");
  __shared__ ChanShared cs;
  __shared__ BlockParams params[2];
  __shared__ uint2 lut[8];
  __shared__ uint32_t tbl[SMEM_TBL];
  __shared__ uint32_t alias_tbl[2][48];  // per ring slot: tbl[0..47] with entry 0 = the block's stale bits (rule A6)
  __shared__ __align__(16) int totals[12];
  __shared__ __align__(8) uint32_t kconst[2];  // {k1, k2048} for the segment loop's volatile reads (track_seg.cuh, SEG_U2)
  __shared__ __align__(8) uint64_t dfull[2], pfull[2], empty[2], tfull;
  extern __shared__ __align__(128) uint8_t tiles[];
  // mixer table behind the two tiles; the segment form ORs the sample code into the table address, so there the
  // table starts on a 2048-byte boundary of the shared window (the launch reserves the slack)
  uint8_t *vlut_raw = tiles + 2 * (size_t)tile_bytes;
  if (SEGH > 0) vlut_raw += (2048u - (smem_u32(vlut_raw) & 2047u)) & 2047u;
  uint32_t *vlut = reinterpret_cast<uint32_t *>(vlut_raw);

  __shared__ int s_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t blk_bytes = bytes_for(fmt, a.nsamp);
  // The control lane sits in the CTA's first warp (measured 1.6 % faster at 64 streams than in the last one: its
  // dependent chain is what the correlator warps wait for, and it is issued sooner there).
  constexpr int CTRL = 0, CTRL_WARP = 0;
  const int ctid = tid - 32;  // index among the correlator threads
  SchedQueue *const wq = a.sched;
  // channel-independent tables first: they fill while the control lane may still be waiting for its item
  fill_lo_lut(lut);
  if (tid == 0) {
    kconst[0] = a.k1;
    kconst[1] = a.k2048;
  }
  if (packed_native) {
    for (int i = tid; i < 128 * 32; i += WS_THREADS) {
      const int e = i >> 5, ph = e >> 4, code = e & 15;
      const int I = sample_val((uint32_t)code), Q = sample_val((uint32_t)code >> 2);
      const int ival = lo_i(ph) * I + lo_q(ph) * Q, qval = lo_q(ph) * I - lo_i(ph) * Q;  // correlator.c:214-215
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
    TCHECK(6, (unsigned)s_item < wq->total);  // the item behind the ticket exists
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
    long long f = code_table_base(tbl_prn) + i;
    tbl[i] = (code_has_fast_row(tbl_prn) && f < TABLE_ENTRIES) ? a.code_table[f] : 0u;
  }
  if (tid < 96) {
    const int t = tid % 48;
    long long f = code_table_base(tbl_prn) + t;
    alias_tbl[tid / 48][t] = (code_has_fast_row(tbl_prn) && f < TABLE_ENTRIES) ? a.code_table[f] : 0u;
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
#ifdef TRACK_CHECK
    // hand-over: bit 2 marks a channel whose slice is running; the CTA that ran the previous slice cleared it with the
    // store of its state, before it pushed this item
    TCHECK(7, (atomicOr(&a.chan_flags[s * NCH + ch], 4) & 4) == 0);
#endif
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
  if (warp == CTRL_WARP) {
    if (lane != 0) return;
    uint32_t inv_of = 0;  // code NCO increment `dinv` belongs to
    double dinv = 0.0;
    auto publish = [&](int slot, const StepParams &sp, bool event) {
      BlockParams &p = params[slot];
      p.cph0 = sp.cph0; p.kph0 = sp.kph0; p.cinc = sp.cinc; p.kinc = sp.kinc;
      p.hc0 = sp.hc0; p.w1 = sp.w1; p.stale_idx = sp.stale_idx; p.stale_bits = sp.stale_bits;
      p.mode = sp.mode; p.event = event ? 1 : 0;
      if constexpr (SEGH > 0) {
        if (sp.mode == MODE_FAST) {
          const uint32_t form = seg_kinc_ok(sp.kinc) ? 1u : ((packed_native && seg_kinc_ok16(sp.kinc)) ? 2u : 0u);
          const bool ok = form != 0;
          if (ok && sp.kinc != inv_of) {  // the division only when the code NCO word changed
            dinv = 1.0 / (double)sp.kinc;
            inv_of = sp.kinc;
          }
          p.wtot = (uint32_t)(((unsigned long long)sp.kph0 + (unsigned long long)a.nsamp * sp.kinc) >> 32);
          p.seg = form;
          p.dinv = dinv;
          p.kseg = 7u * sp.kinc;
        }
      }
      alias_tbl[slot][0] = sp.stale_bits;
      mbar_arrive(&pfull[slot]);  // release: the stores above are visible to whoever observes the phase
    };
    bool event = block_is_event(sp, a, nblocks <= 1);
#ifdef EXP_ALL_QUIET  // timing experiment only (results are wrong): every block handled as a quiet block
    if (sp.mode == MODE_FAST && nblocks > 1) event = false;
#endif
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
#ifdef EXP_ALL_QUIET
        if (sp.mode == MODE_FAST && b + 2 != nblocks) event = false;
#endif
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
#ifdef EXP_EARLY_PUBLISH  // timing experiment only (results are wrong): what the kernel would do without the dump -> ISR -> next block dependency
      bool early = false;
      ChRegs saved_r = cs.r;
      if (was_mode == MODE_FAST && !last) {
        StepParams t = sp;
        prepare_block_regs(cs, t, a, tbl_prn);
        t.stale_bits = t.mode == MODE_FAST ? tbl[t.stale_idx] : 0u;
        publish(nslot, t, block_is_event(t, a, b + 2 == nblocks));
        early = true;
      }
#endif
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
        {  // a copy: only it has its address taken, so the loop's own block parameters stay in registers (with `sp`
           // itself passed, every update of it in this loop is also stored to the stack: ~36 local stores per block)
          const StepParams p_tmp = sp;
          serial_block(cs, p_tmp, a.code_table, fmt, a.nsamp, tiles + (size_t)slot * tile_bytes);
        }
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
#ifdef EXP_EARLY_PUBLISH
      if (early) {
        cs.r.w_carr_hi = saved_r.w_carr_hi; cs.r.w_carr_lo = saved_r.w_carr_lo; cs.r.w_code_hi = saved_r.w_code_hi;
        cs.r.w_code_lo = saved_r.w_code_lo; cs.r.w_slew = saved_r.w_slew;
        cs.halted = 0;
      }
#endif
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
#ifdef EXP_EARLY_PUBLISH
        if (!early)
#endif
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
  const int i0 = ctid * SPT;
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
      if constexpr (SEGH > 0) {
        BlockParams bp;
        bp.cph0 = cph0; bp.kph0 = kph0; bp.cinc = cinc; bp.kinc = kinc;
        bp.hc0 = hc0; bp.w1 = w1; bp.stale_idx = stale_idx; bp.stale_bits = stale_bits;
        const uint2 p3 = reinterpret_cast<const uint2 *>(&params[slot])[5];
        bp.wtot = p3.x;
        bp.seg = p3.y;
        bp.dinv = params[slot].dinv;
        const SampleCtx sc{cph0, kph0, cinc, kinc, hc0, w1, stale_idx, tile, tbl, lut, fmt};
        // what the segment loop may read: the tile plus the 48 bytes idle trailing segments run past it (the other tile or
        // the mixer table follow), the code-table window, the mixer table
        const SegBounds bnd{smem_u32(tile), smem_u32(tile) + (uint32_t)tile_bytes + (packed_native ? 64u : 224u), smem_u32(tbl), smem_u32(tbl) + 4u * SMEM_TBL,
#ifdef TRACK_CHECK_SELFTEST  // a deliberately wrong bound: the check must fire (tools/gpu_evidence.sh)
                            smem_u32(vlut), smem_u32(vlut) + 64u * 32u * 4u};
#else
                            smem_u32(vlut), smem_u32(vlut) + 128u * 32u * 4u};
#endif
        constexpr int SEGH16 = (SEGH * 8 + 15) / 16;  // segments per thread at 15-16 samples each: (CT-1)*SEGH16 covers the ~522 of a block
        bool done = false;
        if constexpr (packed_native) {
          if (bp.seg == 1) {
            seg_block<CT, SEGH>(bp, sc, smem_u32(tile), smem_u32(tbl), smem_u32(alias_tbl[slot]), vlut_lane,
                                PipeK{a.k1, a.k8, a.k128, a.k2048}, a.nsamp, ctid, sumA, sumB, anyB, bnd, smem_u32(&params[slot].kseg),
                                smem_u32(kconst));
            done = true;
          } else if (bp.seg == 2) {
            seg_block<CT, SEGH16, 16>(bp, sc, smem_u32(tile), smem_u32(tbl), smem_u32(alias_tbl[slot]), vlut_lane,
                                      PipeK{a.k1, a.k8, a.k128, a.k2048}, a.nsamp, ctid, sumA, sumB, anyB, bnd);
            done = true;
          }
        } else {
          if (bp.seg == 1) {  // int8 I,Q samples: the LO table instead of the mixer table
            seg_block<CT, SEGH, 8, true>(bp, sc, smem_u32(tile), smem_u32(tbl), smem_u32(alias_tbl[slot]), smem_u32(lut),
                                         PipeK{a.k1, a.k8, a.k128, a.k2048}, a.nsamp, ctid, sumA, sumB, anyB, bnd);
            done = true;
          }
        }
        if (!done) generic_block<CT>(sc, a.nsamp, ctid, sumA, sumB, anyB);
      } else {
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

