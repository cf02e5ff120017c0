// ingest.cu -- streaming ingest: a circular buffer between a sample producer (front-end reader, file
// reader, socket ...) and the GPU channel loop, so a record can be tracked while it is still arriving.
//
// What it replaces (FE = trunk/FRONT_END_PROJECT/PC_SIDE_SOFTWARE/WIN/GPS1A_SAMPLER/src):
//   class CircularBuffer                       FE/CircularBuffer.h:9-193  (positions, totals, overflow rule :126-131,
//                                              contiguous output run :74-85)
//   CollectFromUSB / WriteBufferToFile threads FE/win32_sampler.h:232-364 (producer fills in USB_READ_SIZE pieces,
//                                              consumer drains whole FILE_WRITE_SIZE pieces)
// Here the consumer is not a file writer but gnssb200's tracking kernel: gnssb200_ingest_pump() takes the
// whole 512-us blocks that are available in one contiguous run, copies them from the pinned ring to one of
// two device staging buffers and launches the closed channel loop on them; receiver state stays on the device
// between pumps (the kernel resumes from gnssb200_rx.blocks_done), so the dump records are identical to a
// single pass over the complete record.  The ring size is a whole number of blocks, so a block never wraps.
#include <mutex>
#include <new>
#include <string.h>

#include "common.cuh"

struct gnssb200_ingest {
  gnssb200_handle *h;
  int stream, fmt, nsamp;
  size_t blk_bytes;
  uint8_t *ring;              // pinned host memory
  size_t ring_size;           // CircularBufferSize
  std::mutex lock;            // CheckReadWriteStatus
  size_t in_pos, out_pos;     // InputBufferPos / OutputBufferPos
  unsigned long long tot_loaded, tot_output;  // TotBytesLoaded / TotBytesOutput
  int finished, overflow;     // FinishedFillingBuffer / CircularBufferOverFlow
  uint8_t *d_stage[2];
  long long stage_blocks;
  cudaEvent_t ev_used[2], ev_copied;
  cudaStream_t st;
  int buf;
  gnssb200_dump *d_dumps;
  int32_t *d_cnt;
  int dump_cap;
  long long blocks_done;
};

static size_t ingest_blk_bytes(int fmt, int nsamp) {
  return fmt == GNSSB200_FMT_INT8_IQ ? (size_t)nsamp * 2 : (fmt == GNSSB200_FMT_PACKED2 ? (size_t)nsamp / 2 : (size_t)nsamp);
}

extern "C" gnssb200_ingest *gnssb200_ingest_open(gnssb200_handle *h, int stream, int fmt, int nsamp, int64_t ring_blocks, int dump_cap) {
  if (!h || stream < 0 || stream >= h->n_streams || fmt < 0 || fmt > 2 || nsamp <= 0 || ring_blocks < 2 ||
      (fmt == GNSSB200_FMT_PACKED2 && (nsamp & 1))) {
    gnssb200_set_error(-6, "gnssb200_ingest_open: bad arguments", __FILE__, __LINE__);
    return nullptr;
  }
  if (cudaSetDevice(h->device) != cudaSuccess) return nullptr;
  gnssb200_ingest *g = new (std::nothrow) gnssb200_ingest();
  if (!g) return nullptr;
  g->h = h;
  g->stream = stream;
  g->fmt = fmt;
  g->nsamp = nsamp;
  g->blk_bytes = ingest_blk_bytes(fmt, nsamp);
  g->ring_size = g->blk_bytes * (size_t)ring_blocks;
  g->in_pos = g->out_pos = 0;
  g->tot_loaded = g->tot_output = 0;
  g->finished = g->overflow = 0;
  g->stage_blocks = ring_blocks < 1024 ? ring_blocks : 1024;
  g->buf = 0;
  g->dump_cap = dump_cap;
  g->blocks_done = 0;
  g->ring = nullptr;
  g->d_stage[0] = g->d_stage[1] = nullptr;
  g->d_dumps = nullptr;
  g->d_cnt = nullptr;
  cudaError_t e = cudaHostAlloc(&g->ring, g->ring_size, cudaHostAllocDefault);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) {
    e = cudaMalloc(&g->d_stage[i], g->blk_bytes * (size_t)g->stage_blocks + 256);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_used[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_copied, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking);
  if (e == cudaSuccess && dump_cap > 0) {
    e = cudaMalloc(&g->d_dumps, sizeof(gnssb200_dump) * (size_t)NCH * dump_cap);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_cnt, sizeof(int32_t) * NCH);
    if (e == cudaSuccess) e = cudaMemset(g->d_cnt, 0, sizeof(int32_t) * NCH);
  }
  if (e != cudaSuccess) {
    gnssb200_set_error((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    if (g->ring) cudaFreeHost(g->ring);
    cudaFree(g->d_stage[0]);
    cudaFree(g->d_stage[1]);
    cudaFree(g->d_dumps);
    cudaFree(g->d_cnt);
    delete g;
    return nullptr;
  }
  return g;
}

extern "C" void gnssb200_ingest_close(gnssb200_ingest *g) {
  if (!g) return;
  cudaSetDevice(g->h->device);
  cudaStreamSynchronize(g->st);
  cudaFreeHost(g->ring);
  for (int i = 0; i < 2; i++) {
    cudaFree(g->d_stage[i]);
    cudaEventDestroy(g->ev_used[i]);
  }
  cudaEventDestroy(g->ev_copied);
  cudaStreamDestroy(g->st);
  cudaFree(g->d_dumps);
  cudaFree(g->d_cnt);
  delete g;
}

// Producer side.  All-or-nothing: returns `bytes` when the data was taken, 0 after setting the overflow flag
// when it does not fit (CircularBuffer::AdvanceInputBufferPosition :113-133 flags the overrun and the
// collector stops; here nothing is overwritten).
extern "C" int64_t gnssb200_ingest_write(gnssb200_ingest *g, const void *data, int64_t bytes) {
  if (!g || !data || bytes < 0) return -1;
  size_t pos;
  {
    std::lock_guard<std::mutex> lk(g->lock);
    if (g->tot_loaded + (unsigned long long)bytes > g->tot_output + g->ring_size) {
      g->overflow = 1;
      return 0;
    }
    pos = g->in_pos;
  }
  // single producer: the region [pos, pos+bytes) is free until tot_loaded is advanced below
  const size_t first = (size_t)bytes < g->ring_size - pos ? (size_t)bytes : g->ring_size - pos;
  memcpy(g->ring + pos, data, first);
  if ((size_t)bytes > first) memcpy(g->ring, (const uint8_t *)data + first, (size_t)bytes - first);
  {
    std::lock_guard<std::mutex> lk(g->lock);
    g->in_pos += (size_t)bytes;
    if (g->in_pos >= g->ring_size) g->in_pos -= g->ring_size;
    g->tot_loaded += (unsigned long long)bytes;
  }
  return bytes;
}

extern "C" void gnssb200_ingest_finish(gnssb200_ingest *g) {  // SetFinishedLoadingData
  if (!g) return;
  std::lock_guard<std::mutex> lk(g->lock);
  g->finished = 1;
}

// Consumer side: track the whole blocks available in one contiguous run (at most max_blocks; <= 0: no limit).
// Returns the number of blocks handed to the GPU (0: nothing available yet), < 0 on error.
extern "C" int64_t gnssb200_ingest_pump(gnssb200_ingest *g, int64_t max_blocks) {
  if (!g) return -1;
  size_t pos;
  long long n;
  {
    std::lock_guard<std::mutex> lk(g->lock);  // GetAvailableOutputBlockSize :74-85
    const unsigned long long avail = g->tot_loaded - g->tot_output;
    const size_t contig = g->ring_size - g->out_pos;
    const unsigned long long run = avail < contig ? avail : contig;
    n = (long long)(run / g->blk_bytes);
    pos = g->out_pos;
  }
  if (n > g->stage_blocks) n = g->stage_blocks;
  if (max_blocks > 0 && n > max_blocks) n = max_blocks;
  if (n <= 0) return 0;
  gnssb200_handle *h = g->h;
  cudaError_t e = cudaSetDevice(h->device);
  const int b = g->buf;
  if (e == cudaSuccess) e = cudaStreamWaitEvent(g->st, g->ev_used[b], 0);  // kernels of two pumps ago are done with this buffer
  if (e == cudaSuccess) e = cudaMemcpyAsync(g->d_stage[b], g->ring + pos, g->blk_bytes * (size_t)n, cudaMemcpyHostToDevice, g->st);
  if (e == cudaSuccess) e = cudaEventRecord(g->ev_copied, g->st);
  if (e != cudaSuccess) {
    gnssb200_set_error((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    return -1;
  }
  // kernel indexes samples, dump records and counters by absolute stream number
  const size_t stride = g->blk_bytes * (size_t)g->stage_blocks + 256;
  gnssb200_dump *dumps = g->d_dumps ? g->d_dumps - (size_t)g->stream * NCH * g->dump_cap : nullptr;
  int32_t *cnt = g->d_cnt ? g->d_cnt - (size_t)g->stream * NCH : nullptr;
  int rc = track_launch(h, g->stream, 1, g->d_stage[b], stride, g->fmt, g->nsamp, n, 1, dumps, g->dump_cap, cnt, g->st);
  if (rc) return -1;
  cudaEventRecord(g->ev_used[b], g->st);
  g->buf ^= 1;
  // the ring region may be overwritten by the producer once the DMA has read it
  e = cudaEventSynchronize(g->ev_copied);
  if (e != cudaSuccess) {
    gnssb200_set_error((int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    return -1;
  }
  {
    std::lock_guard<std::mutex> lk(g->lock);  // AdvanceOutputBufferPosition :97-111
    g->out_pos += g->blk_bytes * (size_t)n;
    if (g->out_pos >= g->ring_size) g->out_pos -= g->ring_size;
    g->tot_output += g->blk_bytes * (unsigned long long)n;
  }
  g->blocks_done += n;
  return n;
}

extern "C" int gnssb200_ingest_status(gnssb200_ingest *g, gnssb200_ingest_stat *out) {
  if (!g || !out) return -1;
  std::lock_guard<std::mutex> lk(g->lock);
  out->bytes_loaded = (int64_t)g->tot_loaded;
  out->bytes_output = (int64_t)g->tot_output;
  out->bytes_in_buffer = (int64_t)(g->tot_loaded - g->tot_output);  // DataLeftInBuffer
  out->ring_bytes = (int64_t)g->ring_size;
  out->blocks_done = g->blocks_done;
  out->finished = g->finished;
  out->overflow = g->overflow;
  return 0;
}

// Waits for the kernels issued so far; copies the dump records and per-channel counts to the host
// (either pointer may be NULL).
extern "C" int gnssb200_ingest_sync(gnssb200_ingest *g, gnssb200_dump *h_dumps, int32_t *h_count) {
  if (!g) return -1;
  CUDA_TRY(cudaSetDevice(g->h->device));
  if (g->d_dumps && h_dumps)
    CUDA_TRY(cudaMemcpyAsync(h_dumps, g->d_dumps, sizeof(gnssb200_dump) * (size_t)NCH * g->dump_cap, cudaMemcpyDeviceToHost, g->st));
  if (g->d_cnt && h_count) CUDA_TRY(cudaMemcpyAsync(h_count, g->d_cnt, sizeof(int32_t) * NCH, cudaMemcpyDeviceToHost, g->st));
  CUDA_TRY(cudaStreamSynchronize(g->st));
  return 0;
}
