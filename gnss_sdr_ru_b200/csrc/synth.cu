// synth.cu -- device-side synthetic IF record generator (bench / test input only; not on the hot path).
//
// Same signal model as gnss_sdr_ru_b200/synth.py (which follows the reference's only generator,
// SIM/glonass_l3_generator.sce:60-186): per emitter A*code(t)*data(t)*exp(-i(2*pi*f*t+phi0)),
// phase-continuous integer NCOs, complex AWGN of unit power, 2-bit quantiser (thresholds 0 and
// 1 sigma) -> {-3,-1,+1,+3}.  Written straight into int8 I,Q or the packed 2-bit layout, so the
// 20 GB of the 64-stream configuration never has to cross PCIe.  Bit patterns differ from the
// numpy generator (different RNG); parity checks copy the generated record back to the host.
#include <math.h>

#include <vector>

#include "common.cuh"

struct SynthSat {
  uint64_t carr_ph0, carr_inc;   // cycles, 64-bit fraction
  uint64_t code_ph0, code_inc;   // chips, 32.32 fixed point (mod code_len)
  float amp;
  int32_t code_len;
  int32_t table_row;             // row in the chip table
  uint32_t data_seed;            // 0: no data
  uint32_t samples_per_bit;
  uint32_t bits_off, n_bits;     // explicit data bits: n_bits entries of the bit pool starting at bits_off (n_bits = 0: none)
};

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27; z *= 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return z;
}

// chips: int8 [rows][1024]
__global__ void synth_kernel(uint8_t *out, size_t stride, int fmt, int64_t n_samples, const SynthSat *sats, int n_sats,
                             const int8_t *chips, uint64_t seed, const uint8_t *bit_pool) {
  const int s = blockIdx.y;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;  // 4 complex samples per thread
  if (i0 >= n_samples) return;
  float accI[4] = {0, 0, 0, 0}, accQ[4] = {0, 0, 0, 0};
  for (int k = 0; k < n_sats; k++) {
    const SynthSat st = sats[s * n_sats + k];
    if (st.amp == 0.f) continue;
    const uint64_t period = (uint64_t)st.code_len << 32;
    // code phase at i0: (ph0 + i0*inc) mod period, 128-bit safe because i0*inc < 2^63 for n < 2^31
    uint64_t cp = (st.code_ph0 + (uint64_t)i0 * st.code_inc) % period;
    uint64_t ph = st.carr_ph0 + (uint64_t)i0 * st.carr_inc;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int chip = chips[st.table_row * 1024 + (int)(cp >> 32)];
      float a = st.amp * (float)chip;
      if (st.n_bits) {
        const uint64_t bit = (uint64_t)(i0 + e) / st.samples_per_bit;
        if (bit_pool[st.bits_off + (uint32_t)(bit % st.n_bits)]) a = -a;
      } else if (st.data_seed) {
        const uint64_t bit = (uint64_t)(i0 + e) / st.samples_per_bit;
        if (mix64(bit * 0x9E3779B97F4A7C15ULL + st.data_seed) & 1) a = -a;
      }
      float sn, cs;
      sincospif((float)(int32_t)(ph >> 32) * (1.0f / 2147483648.0f), &sn, &cs);  // angle = 2*pi*frac
      accI[e] += a * cs;
      accQ[e] -= a * sn;
      ph += st.carr_inc;
      cp += st.code_inc;
      if (cp >= period) cp -= period;
    }
  }
  uint32_t codes = 0;
  int8_t vals[8];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const uint64_t r = mix64(seed + ((uint64_t)s << 40) + (uint64_t)(i0 + e));
    const float u1 = ((float)(uint32_t)(r >> 32) + 1.0f) * (1.0f / 4294967296.0f);
    const float u2 = (float)(uint32_t)r * (1.0f / 4294967296.0f);
    const float rad = sqrtf(-logf(u1));  // sigma^2 = 1/2 per component: sqrt(-2 ln u * 1/2)
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    const float vi = accI[e] + rad * cs, vq = accQ[e] + rad * sn;
    const float thr = 0.70710678f;
    const int ci = (vi < 0.f ? 1 : 0) | (fabsf(vi) > thr ? 2 : 0);
    const int cq = (vq < 0.f ? 1 : 0) | (fabsf(vq) > thr ? 2 : 0);
    codes |= (uint32_t)(ci | (cq << 2)) << (4 * e);
    vals[2 * e] = (int8_t)((ci & 2 ? 3 : 1) * (ci & 1 ? -1 : 1));
    vals[2 * e + 1] = (int8_t)((cq & 2 ? 3 : 1) * (cq & 1 ? -1 : 1));
  }
  uint8_t *base = out + (size_t)s * stride;
  if (fmt == GNSSB200_FMT_PACKED2) {
    *reinterpret_cast<uint16_t *>(base + (i0 >> 1)) = (uint16_t)codes;
  } else {
    uint2 v;
    v.x = (uint8_t)vals[0] | ((uint32_t)(uint8_t)vals[1] << 8) | ((uint32_t)(uint8_t)vals[2] << 16) | ((uint32_t)(uint8_t)vals[3] << 24);
    v.y = (uint8_t)vals[4] | ((uint32_t)(uint8_t)vals[5] << 8) | ((uint32_t)(uint8_t)vals[6] << 16) | ((uint32_t)(uint8_t)vals[7] << 24);
    *reinterpret_cast<uint2 *>(base + 2 * i0) = v;
  }
}

// chip table of a handle: rows 1..32 GPS C/A (same generator as the correlator table), row 0 GLONASS ST code
int chip_table(gnssb200_handle *h, const int8_t **out) {
  if (!h->d_chips) {
    std::vector<int8_t> chips(33 * 1024, 0);
    std::vector<uint32_t> table(TABLE_ENTRIES + 1);
    build_code_table_host(table.data());
    for (int prn = 1; prn <= 32; prn++)
      for (int c = 0; c < 1023; c++) chips[prn * 1024 + c] = (int8_t)(table[prn * HALF_CHIPS + 2 * c] & 0xff);  // early[2c] = chip c
    {  // GLONASS ST: 9-stage register, output stage 7, feedback 5^9 (generateSTcode.sci:35-42)
      int reg[9];
      for (int i = 0; i < 9; i++) reg[i] = 1;
      for (int c = 0; c < 511; c++) {
        chips[c] = (int8_t)(2 * reg[6] - 1);
        const int fb = reg[4] ^ reg[8];
        for (int i = 8; i > 0; i--) reg[i] = reg[i - 1];
        reg[0] = fb;
      }
    }
    int8_t *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, chips.size()));
    cudaError_t e = cudaMemcpy(d, chips.data(), chips.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(d);
      CUDA_TRY(e);
    }
    h->d_chips = d;
  }
  *out = h->d_chips;
  return 0;
}

extern "C" int gnssb200_synth(gnssb200_handle *h, void *d_out, size_t stride, int fmt, int n_streams, int64_t n_samples,
                              const gnssb200_synth_sat *sats, int n_sats, uint64_t seed, void *cuda_stream) {
  if (!h || n_streams <= 0 || n_samples <= 0 || (n_samples & 3) || fmt == GNSSB200_FMT_INT8_I) {
    gnssb200_set_error(-6, "gnssb200_synth: bad arguments (n_samples must be a multiple of 4)", __FILE__, __LINE__);
    return -6;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int8_t *d_chips = nullptr;
  if (int rc = chip_table(h, &d_chips)) return rc;
  std::vector<SynthSat> hs((size_t)n_streams * n_sats);
  std::vector<uint8_t> pool;  // explicit data bits of all emitters, back to back
  for (size_t i = 0; i < hs.size(); i++) {
    const gnssb200_synth_sat &g = sats[i];
    SynthSat &d = hs[i];
    const bool glo = g.system == GNSSB200_SYS_GLONASS;
    d.code_len = glo ? 511 : 1023;
    d.table_row = glo ? 0 : g.prn;
    d.amp = (g.cn0_dbhz <= 0.0) ? 0.f : (float)sqrt(pow(10.0, g.cn0_dbhz / 10.0) / g.samp_rate);
    auto frac64 = [](double x) {  // fractional part as a 64-bit fraction
      x -= floor(x);
      return (uint64_t)(x * 18446744073709551616.0);
    };
    d.carr_inc = frac64(g.carrier_hz / g.samp_rate);
    d.carr_ph0 = frac64(g.carrier_phase_cycles);
    d.code_inc = (uint64_t)(g.code_hz / g.samp_rate * 4294967296.0);
    d.code_ph0 = (uint64_t)(fmod(g.code_phase_chips, (double)d.code_len) * 4294967296.0);
    d.data_seed = (uint32_t)g.data_seed;
    d.samples_per_bit = (uint32_t)(g.samp_rate / (g.data_rate_hz > 0 ? g.data_rate_hz : 50.0));
    d.bits_off = d.n_bits = 0;
    if (g.data_bits && g.n_data_bits > 0) {
      d.bits_off = (uint32_t)pool.size();
      d.n_bits = (uint32_t)g.n_data_bits;
      pool.insert(pool.end(), g.data_bits, g.data_bits + g.n_data_bits);
    }
  }
  uint8_t *d_pool = nullptr;
  if (!pool.empty()) {
    CUDA_TRY(cudaMalloc(&d_pool, pool.size()));
    CUDA_TRY(cudaMemcpyAsync(d_pool, pool.data(), pool.size(), cudaMemcpyHostToDevice, st));
  }
  SynthSat *d_sats = nullptr;
  CUDA_TRY(cudaMalloc(&d_sats, hs.size() * sizeof(SynthSat)));
  CUDA_TRY(cudaMemcpyAsync(d_sats, hs.data(), hs.size() * sizeof(SynthSat), cudaMemcpyHostToDevice, st));
  const int threads = 256;
  dim3 grid((unsigned)((n_samples / 4 + threads - 1) / threads), (unsigned)n_streams);
  synth_kernel<<<grid, threads, 0, st>>>((uint8_t *)d_out, stride, fmt, n_samples, d_sats, n_sats, d_chips, seed, d_pool);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));  // hs / d_sats lifetimes
  cudaFree(d_sats);
  cudaFree(d_pool);
  return 0;
}
