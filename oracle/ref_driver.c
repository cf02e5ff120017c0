/* TEST INFRASTRUCTURE ONLY.  A loop around the reference's own Sim_GP2021_int + gpsisr
 * (the body of main(), OSG/osgnss_next_step.c:168-184, without display) that logs one
 * gnssb200_dump per correlator dump natively, so that timing and parity runs do not pay Python
 * overhead per block.  Compiled together with the reference objects by build_ref.sh; everything
 * it calls is the reference's code. */
#include <stdint.h>

#include "../include/gnssb200.h"
#include "globals.h" /* the reference's header (via -I), for tracking_channel / chan[] */

extern int REG_read[256], REG_write[256];
extern void Sim_GP2021_int(char *IF, long nsamp);
extern void gpsisr(void);

long ref_run(const char *IF, long nsamp, long nblocks, long block0, gnssb200_dump *dumps, int cap, int32_t *count) {
  long b;
  for (b = 0; b < nblocks; b++) {
    Sim_GP2021_int((char *)IF + 2 * nsamp * b, nsamp);
    int status = REG_read[0x82];
    gpsisr();
    if (!dumps) continue;
    for (int ch = 0; ch < N_CHANNELS; ch++) {
      if (!(status & (1 << ch)) || count[ch] >= cap) continue;
      gnssb200_dump *d = &dumps[(long)ch * cap + count[ch]++];
      d->block = (int32_t)(block0 + b);
      d->ch = (int16_t)ch;
      d->state = (int16_t)chan[ch].state;
      for (int a = 0; a < 6; a++) d->acc[a] = REG_read[(ch << 3) + 0x84 + a];
      d->carrier_incr = (uint32_t)((REG_write[(ch << 3) + 3] << 16) + REG_write[(ch << 3) + 4]);
      d->code_incr = (uint32_t)((REG_write[(ch << 3) + 5] << 16) + REG_write[(ch << 3) + 6]);
      d->n_freq = (int16_t)chan[ch].n_freq;
      d->codes = (int16_t)chan[ch].codes;
      d->slew = REG_write[(ch << 3) + 0x84];
    }
  }
  return b;
}
