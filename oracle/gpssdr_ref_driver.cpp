// TEST INFRASTRUCTURE ONLY -- C entry points around the reference's own GPS-SDR primitives, compiled in
// place from RT/objects/fft.cpp (-DNO_SIMD: its portable butterflies), RT/simd/x86.cpp and
// RT/accessories/misc.cpp by oracle/build_ref_gpssdr.sh.  Used to pin oracle/gpssdr_oracle.c.
#include "includes.h"
#include "fft.h"

extern "C" {
void gsr_fft(CPX *x, int n, int *R, int inverse, int shuf) {
  FFT f(n, R);
  if (inverse)
    f.doiFFT(x, shuf != 0);
  else
    f.doFFT(x, shuf != 0);
}
void gsr_cmulsc(CPX *A, CPX *B, CPX *C, int cnt, int shift) { x86_cmulsc(A, B, C, cnt, shift); }
void gsr_cmuls(CPX *A, CPX *B, int cnt, int shift) { x86_cmuls(A, B, cnt, shift); }
void gsr_cacc(CPX *A, MIX *B, int cnt, int *ia, int *qa) { x86_cacc(A, B, cnt, ia, qa); }
void gsr_cmag(CPX *A, int cnt) { x86_cmag(A, cnt); }
void gsr_max(int *A, int *index, int *magt, int cnt) { x86_max(A, index, magt, cnt); }
void gsr_sine_gen(CPX *d, double f, double fs, int n) { sine_gen(d, f, fs, n); }
void gsr_wipeoff_gen(MIX *d, double f, double fs, int n) { wipeoff_gen(d, f, fs, n); }
}
