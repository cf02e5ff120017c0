// track_loop.cuh -- track_loop_kernel: the generic, barrier-synchronised channel loop (any block length, unaligned /
// ragged / I-only records; GNSSB200_TRACK_WS=0 selects it for A/B runs).  One CTA per (stream, channel) for the whole run.
#pragma once
#include "track_common.cuh"

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
    long long f = code_table_base(tbl_prn) + i;
    tbl[i] = (code_has_fast_row(tbl_prn) && f < TABLE_ENTRIES) ? a.code_table[f] : 0u;
  }
  if (packed_native) {
    for (int i = tid; i < 128 * 32; i += blockDim.x) {
      const int e = i >> 5, ph = e >> 4, code = e & 15;
      const int I = sample_val((uint32_t)code), Q = sample_val((uint32_t)code >> 2);
      const int ival = lo_i(ph) * I + lo_q(ph) * Q, qval = lo_q(ph) * I - lo_i(ph) * Q;  // correlator.c:214-215
      vlut[i] = (uint32_t)(ival + 65536 * qval);
    }
  }
  if (tid < 48) {
    long long f = code_table_base(tbl_prn) + tid;
    alias_tbl[tid] = (code_has_fast_row(tbl_prn) && f < TABLE_ENTRIES) ? a.code_table[f] : 0u;
  }
  for (int i = tid; i < 256; i += blockDim.x) {
    uint32_t wv = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) wv |= (uint32_t)(sample_val((uint32_t)i >> (2 * e)) & 0xff) << (8 * e);
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
      const long long f = code_table_base(tbl_prn) + sp_s.stale_idx;
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
  long long prefetched = -1;  // last block whose TMA load was issued (uniform over the CTA)
  long long b = 0;
  for (; b < a.nblocks; b++) {
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
      if (b + 1 < a.nblocks) prefetched = b + 1;
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
        else if (sp.mode == MODE_SERIAL) {
          const StepParams p_tmp = sp;  // a copy: `sp` itself must not have its address taken (see track_ws.cuh)
          serial_block(cs, p_tmp, a.code_table, fmt, a.nsamp, use_tma ? tile : blk);
        }
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

  // the loop left early (halt) with the next block's load still in flight: it must land before the CTA exits
  if (use_tma && b < a.nblocks && prefetched == b) mbar_wait(&mbar[b & 1], (uint32_t)((b >> 1) & 1));
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

