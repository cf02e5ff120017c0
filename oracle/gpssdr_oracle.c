/* TEST INFRASTRUCTURE ONLY (oracle) -- plain C restatement of the GPS-SDR fixed-point FFT acquisition.
 *
 * RT = trunk/GNSS_SOFTWARE_RECEIVERS/REALTIME_RECEIVERS/GPS/GPS_SDR_REAL_TIME_GPS_RECEIVER of the reference.
 * Follows, in the reference's portable (non-SSE) arithmetic:
 *   RT/objects/fft.cpp:56-310      FFT class: twiddles floor(16384*cos/sin), bit-reversal shuffle, rank loops
 *   RT/objects/fft.cpp:312-489     NO_SIMD butterflies: bfly (>>1 pre-scale), bfly_noscale, (x+8192)>>14 rounding
 *   RT/simd/x86.cpp:143-296        x86_cmuls, x86_cmulsc (rounded >>shift), x86_cacc, x86_cmag, x86_max
 *   RT/accessories/misc.cpp:95-166 sine_gen (float phase accumulator), wipeoff_gen (double phase)
 *   RT/objects/acquisition.cpp:68-141   constructor: wipe-off tables, DFT rows, FFT rank-scaling patterns R1/R2
 *   RT/objects/acquisition.cpp:182-236  doPrepIF: 250/500/750 Hz offsets, mix to baseband, forward FFTs, padded rows
 *   RT/objects/acquisition.cpp:244-302  doAcqStrong
 *   RT/objects/acquisition.cpp:433-570  doAcqWeak: 10 ms coherent, 10-point post-correlation DFT (25 Hz), 15
 *                                       non-coherent sums with code-Doppler shift, even / odd 10-ms alignment
 * The acquisition object calls the SSE forms (sse_cmulsc, sse_cacc) on 32-bit x86; the in-tree portable forms
 * restated here are what the reference itself ships for other builds (fft.cpp NO_SIMD, x86.cpp).
 * PINNED at the primitive level: tests/test_gpssdr_oracle.py compares every primitive below with the
 * reference's own fft.cpp (-DNO_SIMD) / x86.cpp / misc.cpp compiled in place (oracle/_ref/libgpssdr_ref.so).
 * The Acquisition class itself cannot be built (USRP headers), so the two pipeline functions are restated only.
 * Only tests/, smoke() and bench.py's CPU leg may use this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int16_t i, q; } CPX;
typedef struct { int16_t i, nq, q, ni; } MIX;

#define SAMPS_MS 2048
#define SAMPLE_FREQUENCY 2048000
#define L1_HZ 1.57542e9
#define TWO_PI 6.283185307179586

/* ---- FFT (fft.cpp) ---- */
typedef struct {
  int N, M;
  int R[16];
  MIX *W, *iW;
  int *BR;
  int32_t *BRX;
} FFT;

static void fft_init(FFT *f, int N, const int *R) {
  const double pi = 3.14159265358979323846264338327;
  f->N = N;
  f->M = 0;
  for (int i = 0; i < 16; i++) f->R[i] = R ? R[i] : 1;
  for (int n = N; n > 1; n >>= 1) f->M++;
  f->W = (MIX *)malloc(sizeof(MIX) * N / 2);
  f->iW = (MIX *)malloc(sizeof(MIX) * N / 2);
  f->BR = (int *)malloc(sizeof(int) * N);
  f->BRX = (int32_t *)malloc(sizeof(int32_t) * N);
  for (int l = 0; l < N / 2; l++) { /* initW :121-149 */
    double phase = (-2 * pi * l) / N;
    double c = floor(16384 * cos(phase)), s = floor(16384 * sin(phase));
    f->W[l].i = (short)c;  f->W[l].q = (short)s;   f->W[l].nq = (short)(-s); f->W[l].ni = (short)c;
    f->iW[l].i = (short)c; f->iW[l].q = (short)(-s); f->iW[l].nq = (short)s; f->iW[l].ni = (short)c;
  }
  for (int l = 0; l < N; l++) { /* initBR :153-171 */
    int index = 0;
    for (int b = 0; b < f->M; b++) {
      index += (l >> b) & 1;
      index <<= 1;
    }
    f->BR[l] = index >> 1;
  }
}
static void fft_free(FFT *f) { free(f->W); free(f->iW); free(f->BR); free(f->BRX); }

static void bfly(CPX *A, CPX *B, const MIX *W, int scale) { /* :401-440 */
  int32_t bi, bq;
  if (scale) {
    A->i >>= 1; A->q >>= 1; B->i >>= 1; B->q >>= 1;
  }
  bi = B->i * W->i - B->q * W->q;
  bq = B->i * W->q + B->q * W->i;
  bi = (bi + 8192) >> 14;
  bq = (bq + 8192) >> 14;
  B->i = A->i - (int16_t)bi;
  B->q = A->q - (int16_t)bq;
  A->i += (int16_t)bi;
  A->q += (int16_t)bq;
}
static void fft_run(FFT *f, CPX *x, int inverse, int shuf) { /* doFFT :173-201 / doiFFT :204-233 */
  if (shuf) { /* doShuffle :298-310 */
    int32_t *p = (int32_t *)x;
    memcpy(f->BRX, p, sizeof(int32_t) * f->N);
    for (int l = 0; l < f->N; l++) p[l] = f->BRX[f->BR[l]];
  }
  int bsize = 1, nblocks = f->N >> 1;
  const MIX *Wt = inverse ? f->iW : f->W;
  for (int r = 0; r < f->M; r++) {
    CPX *a = x, *b = x + bsize;
    for (int blk = 0; blk < nblocks; blk++) { /* rank / rank_noscale :314-357 */
      const MIX *w = Wt;
      for (int j = 0; j < bsize; j++) {
        bfly(a, b, w, f->R[r]);
        a++; b++; w += nblocks;
      }
      a += bsize;
      b += bsize;
    }
    bsize <<= 1;
    nblocks >>= 1;
  }
}

/* ---- x86.cpp ---- */
static void cmulsc(const CPX *A, const CPX *B, CPX *C, int cnt, int shift) { /* x86_cmulsc :181-216, x86_cmuls with C == A */
  const int32_t round = 1 << (shift - 1);
  for (int l = 0; l < cnt; l++) {
    int32_t ai = A[l].i, aq = A[l].q, bi = B[l].i, bq = B[l].q;
    int32_t ti = ai * bi - aq * bq, tq = ai * bq + aq * bi;
    ti += round; tq += round;
    ti >>= shift; tq >>= shift;
    C[l].i = (int16_t)ti;
    C[l].q = (int16_t)tq;
  }
}
static void cacc(const CPX *A, const MIX *B, int cnt, int32_t *ia, int32_t *qa) { /* x86_cacc :220-251 */
  int32_t iacc = 0, qacc = 0;
  for (int l = 0; l < cnt; l++) {
    int32_t ai = A[l].i, aq = A[l].q;
    iacc += ai * B[l].i + aq * B[l].nq;
    qacc += ai * B[l].q + aq * B[l].ni;
  }
  *ia = iacc;
  *qa = qacc;
}
static void cmag(CPX *A, int cnt) { /* x86_cmag :255-269 (in place, int32 over the CPX) */
  int32_t *p = (int32_t *)A;
  for (int l = 0; l < cnt; l++) p[l] = A[l].i * A[l].i + A[l].q * A[l].q;
}
static void imax(const int32_t *A, int32_t *index, int32_t *magt, int cnt) { /* x86_max :273-294 */
  int32_t mag = 0, idx = 0;
  for (int l = 0; l < cnt; l++)
    if (A[l] > mag) { idx = l; mag = A[l]; }
  *index = idx;
  *magt = mag;
}

/* ---- misc.cpp ---- */
static void sine_gen_f(CPX *dest, double f, double fs, int samps) { /* :95-114: FLOAT phase accumulator, cos/sin of a float */
  float phase = 0, phase_step = (float)TWO_PI * f / fs;
  for (int l = 0; l < samps; l++) {
    dest[l].i = (int16_t)floor(16383.0 * cosf(phase));
    dest[l].q = (int16_t)floor(16383.0 * sinf(phase));
    phase += phase_step;
  }
}
static void wipeoff_gen(MIX *dest, double f, double fs, int samps) { /* :148-166 */
  double phase = 0, phase_step = (double)TWO_PI * f / fs;
  for (int l = 0; l < samps; l++) {
    int16_t c = (int16_t)floor(16383.0 * cos(phase)), s = (int16_t)floor(16383.0 * sin(phase));
    dest[l].i = dest[l].ni = c;
    dest[l].q = s;
    dest[l].nq = -s;
    phase += phase_step;
  }
}

/* primitive entry points for the pinning test */
void gso_fft(CPX *x, int n, const int *R, int inverse, int shuf) {
  FFT f;
  fft_init(&f, n, R);
  fft_run(&f, x, inverse, shuf);
  fft_free(&f);
}
void gso_cmulsc(const CPX *A, const CPX *B, CPX *C, int cnt, int shift) { cmulsc(A, B, C, cnt, shift); }
void gso_cacc(const CPX *A, const MIX *B, int cnt, int32_t *ia, int32_t *qa) { cacc(A, B, cnt, ia, qa); }
void gso_cmag(CPX *A, int cnt) { cmag(A, cnt); }
void gso_max(const int32_t *A, int32_t *index, int32_t *magt, int cnt) { imax(A, index, magt, cnt); }
void gso_sine_gen(CPX *d, double f, double fs, int n) { sine_gen_f(d, f, fs, n); }
void gso_wipeoff_gen(MIX *d, double f, double fs, int n) { wipeoff_gen(d, f, fs, n); }

/* ---- Acquisition (acquisition.cpp) ---- */
typedef struct {
  double fif;
  CPX *baseband;        /* [4*310*2048] */
  CPX *baseband_shift;  /* [4*310][2048+201] */
  CPX *wipe[4];         /* 0 / 250 / 500 / 750 Hz, [310*2048] */
  MIX dft[10][10];
  FFT fwd, inv;
  int ms;               /* rows prepared per offset */
} GsoAcq;

#define ROWLEN (SAMPS_MS + 201)

GsoAcq *gso_acq_new(double fif) { /* constructor :68-141 */
  static const int R1[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  static const int R2[16] = {0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 1};
  GsoAcq *a = (GsoAcq *)calloc(1, sizeof *a);
  a->fif = fif;
  a->baseband = (CPX *)calloc((size_t)4 * 310 * SAMPS_MS, sizeof(CPX));
  a->baseband_shift = (CPX *)calloc((size_t)4 * 310 * ROWLEN, sizeof(CPX));
  for (int k = 0; k < 4; k++) {
    a->wipe[k] = (CPX *)calloc((size_t)310 * SAMPS_MS, sizeof(CPX));
    sine_gen_f(a->wipe[k], -fif - 250.0 * k, SAMPLE_FREQUENCY, 10 * SAMPS_MS);
    for (int l = 1; l < 31; l++) memcpy(&a->wipe[k][(size_t)l * 10 * SAMPS_MS], a->wipe[k], sizeof(CPX) * 10 * SAMPS_MS);
  }
  for (int l = 0; l < 10; l++) wipeoff_gen(a->dft[l], (float)l * 25.0 - 112.5, 1000.0, 10);
  fft_init(&a->fwd, SAMPS_MS, R1);
  fft_init(&a->inv, SAMPS_MS, R2);
  return a;
}
void gso_acq_free(GsoAcq *a) {
  free(a->baseband); free(a->baseband_shift);
  for (int k = 0; k < 4; k++) free(a->wipe[k]);
  fft_free(&a->fwd); fft_free(&a->inv);
  free(a);
}
static CPX *row_ptr(GsoAcq *a, int row) { return a->baseband_shift + (size_t)row * ROWLEN; }

void gso_prep_if(GsoAcq *a, int type, const CPX *buff) { /* doPrepIF :182-236 */
  const int ms = type == 0 ? 1 : (type == 1 ? 10 : (type == 2 ? 310 : 1));
  const size_t n = (size_t)ms * SAMPS_MS;
  a->ms = ms;
  memcpy(a->baseband, buff, n * sizeof(CPX));
  cmulsc(a->baseband, a->wipe[1], a->baseband + n, (int)n, 14);
  cmulsc(a->baseband, a->wipe[2], a->baseband + 2 * n, (int)n, 14);
  cmulsc(a->baseband, a->wipe[3], a->baseband + 3 * n, (int)n, 14);
  cmulsc(a->baseband, a->wipe[0], a->baseband, (int)n, 14); /* sse_cmuls, in place */
  for (int l = 0; l < 4 * ms; l++) fft_run(&a->fwd, a->baseband + (size_t)l * SAMPS_MS, 0, 1);
  for (int l = 0; l < 4 * ms; l++) {
    CPX *p = row_ptr(a, l);
    memcpy(p, a->baseband + (size_t)(l + 1) * SAMPS_MS - 100, 100 * sizeof(CPX));
    memcpy(p + 100, a->baseband + (size_t)l * SAMPS_MS, SAMPS_MS * sizeof(CPX));
    memcpy(p + 100 + SAMPS_MS, a->baseband + (size_t)l * SAMPS_MS, 100 * sizeof(CPX));
  }
}

typedef struct { int32_t code_phase; int32_t doppler; uint32_t magnitude; int32_t pad_; } GsoResult; /* Acq_Command_S: code_phase, doppler (int32), magnitude (uint32), RT/includes/structs.h:155-158 */

void gso_acq_strong(GsoAcq *a, const CPX *code, int doppmin, int doppmax, GsoResult *res) { /* doAcqStrong :244-302 */
  int32_t mag = 0, magt = 0, indext = 0;
  CPX msbuff[SAMPS_MS];
  for (int l = doppmin / 1000; l < doppmax / 1000; l++)
    for (int l2 = 0; l2 < 4; l2++) {
      cmulsc(row_ptr(a, l2) + 100 + l, code, msbuff, SAMPS_MS, 10); /* baseband_rows[lcv2]: the four offsets of a 1-ms prep (type 0) */
      fft_run(&a->inv, msbuff, 1, 1);
      cmag(msbuff, SAMPS_MS);
      imax((int32_t *)msbuff, &indext, &magt, SAMPS_MS);
      if (magt > mag) {
        mag = magt;
        res->code_phase = 2048 - indext;
        res->doppler = (int32_t)((l * 1000) + (float)l2 * 250);
        res->magnitude = (uint32_t)mag;
      }
    }
}

void gso_acq_weak(GsoAcq *a, const CPX *code, int doppmin, int doppmax, GsoResult *res) { /* doAcqWeak :433-570 */
  int32_t mag = 0, magt = 0, indext = 0;
  CPX *coherent = (CPX *)malloc(sizeof(CPX) * 10 * SAMPS_MS);
  int32_t *power = (int32_t *)malloc(sizeof(int32_t) * 10 * SAMPS_MS);
  for (int l = doppmin / 1000; l < doppmax / 1000; l++)
    for (int l2 = 0; l2 < 4; l2++)
      for (int k = 0; k < 2; k++) {
        memset(power, 0, sizeof(int32_t) * 10 * SAMPS_MS);
        for (int i = 0; i < 15; i++) {
          for (int l3 = 0; l3 < 10; l3++) {
            cmulsc(row_ptr(a, l2 * 310 + l3 + i * 20 + k * 10) + 100 + l, code, coherent + (size_t)l3 * SAMPS_MS, SAMPS_MS, 9);
            fft_run(&a->inv, coherent + (size_t)l3 * SAMPS_MS, 1, 1);
          }
          const double doppler = (double)(l * 1000) + (float)(l2 * 250);
          const double code_doppler = (double)i * .02 * SAMPLE_FREQUENCY * doppler / L1_HZ;
          const int32_t shift = (int32_t)floor(code_doppler);
          for (int l3 = 0; l3 < SAMPS_MS; l3++) {
            CPX data[10], temp[10];
            for (int j = 0; j < 10; j++) data[j] = coherent[(size_t)j * SAMPS_MS + l3];
            for (int r = 0; r < 10; r++) {
              int32_t ia, qa;
              cacc(data, a->dft[r], 10, &ia, &qa);
              temp[r].i = (int16_t)(ia >> 16);
              temp[r].q = (int16_t)(qa >> 16);
            }
            cmag(temp, 10);
            const int32_t *dt = (const int32_t *)temp;
            const int col = (l3 + shift + SAMPS_MS) % SAMPS_MS;
            for (int r = 0; r < 10; r++) power[(size_t)r * SAMPS_MS + col] += dt[r];
          }
        }
        imax(power, &indext, &magt, 10 * SAMPS_MS);
        if (magt > mag) {
          mag = magt;
          res->code_phase = indext % SAMPS_MS;
          res->doppler = (int32_t)((l * 1000) + (l2 * 250) + (indext / SAMPS_MS) * 25.0);
          res->magnitude = (uint32_t)mag;
        }
      }
  free(coherent);
  free(power);
}

/* doAcqMedium :309-425.  Rows lcv2*20 + lcv3 are read although a 10-ms preparation (type 1) fills rows
 * offset*10 + ms only (:186-194,226-234): lcv2 = 0 reads the 0 Hz rows, lcv2 = 1 the 500 Hz rows (reported as
 * +250 Hz), lcv2 = 2 and 3 read rows 40-49 and 60-69, which hold whatever an EARLIER preparation left there
 * (this object: zeros after gso_acq_new, like a zeroed allocation; rows of the last 310-ms preparation
 * otherwise).  The Doppler loop is inclusive (lcv <= doppmax/1000), the multiply shifts by 10. */
void gso_acq_medium(GsoAcq *a, const CPX *code, int doppmin, int doppmax, GsoResult *res) {
  int32_t mag = 0, magt = 0, indext = 0;
  CPX *coherent = (CPX *)malloc(sizeof(CPX) * 10 * SAMPS_MS);
  CPX *power = (CPX *)malloc(sizeof(CPX) * 10 * SAMPS_MS);
  for (int l = doppmin / 1000; l <= doppmax / 1000; l++)
    for (int l2 = 0; l2 < 4; l2++) {
      for (int l3 = 0; l3 < 10; l3++) {
        cmulsc(row_ptr(a, l2 * 20 + l3) + 100 + l, code, coherent + (size_t)l3 * SAMPS_MS, SAMPS_MS, 10);
        fft_run(&a->inv, coherent + (size_t)l3 * SAMPS_MS, 1, 1);
      }
      for (int l3 = 0; l3 < SAMPS_MS; l3++) {
        CPX data[10];
        for (int j = 0; j < 10; j++) data[j] = coherent[(size_t)j * SAMPS_MS + l3];
        for (int r = 0; r < 10; r++) {
          int32_t ia, qa;
          cacc(data, a->dft[r], 10, &ia, &qa);
          power[(size_t)r * SAMPS_MS + l3].i = (int16_t)(ia >> 16);
          power[(size_t)r * SAMPS_MS + l3].q = (int16_t)(qa >> 16);
        }
      }
      cmag(power, 10 * SAMPS_MS);
      imax((int32_t *)power, &indext, &magt, 10 * SAMPS_MS);
      if (magt > mag) {
        mag = magt;
        res->code_phase = indext % SAMPS_MS;
        res->doppler = (int32_t)((l * 1000) + (l2 * 250) + (indext / SAMPS_MS) * 25.0);
        res->magnitude = (uint32_t)mag;
      }
    }
  free(coherent);
  free(power);
}
