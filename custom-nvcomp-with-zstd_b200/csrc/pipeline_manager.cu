// pipeline_manager.cu -- PipelinedBatchManager: host-resident data streamed through the batch codec.
//
// Replaces src/pipeline_manager.cu of the reference (three host threads handing slots over through
// queues, each batch compressed by the blocking single-buffer call, :101-234).  Here one host thread
// drives three streams; a slot moves through
//     fill (host callback) -> H2D [stream 0] -> compress_async_no_sync [stream 1] -> D2H [stream 2] -> output callback
// and the stages of neighbouring batches overlap because nothing on the way blocks the host except
// the two points where it needs a value: the frame size before the D2H can be issued, and the D2H
// itself before the callback may read the pinned buffer.  Both waits are placed AFTER the next
// batch has been filled and uploaded.
#include "../../include/pipeline_manager.hpp"

#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <new>

namespace cuda_zstd {

namespace {
// the 16-byte result mailbox {bytes, status} of a slot sits behind its pinned output buffer
inline unsigned long long *mailbox(const RingBufferSlot &s) {
  return reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(s.h_output) + s.output_capacity);
}
inline ZstdBatchManager *batch_of(ZstdManager *m) { return static_cast<ZstdBatchManager *>(m); }    // create_manager() builds nothing else
}  // namespace

PipelinedBatchManager::PipelinedBatchManager(const CompressionConfig &config, size_t batch_size_bytes, int num_slots)
    : manager_(create_manager(config)), config_(config), batch_size_(std::max<size_t>(batch_size_bytes, 1)),
      num_slots_(std::max(num_slots, 2)) {
  ring_buffer_.resize((size_t)num_slots_);
  // [0] upload, [1] compress, [2] download as in the reference, then one more compress stream per further slot: a batch
  // of B blocks occupies only B of the encoder's ~3,500 resident warps, so batches in different slots run side by side
  streams_.assign(3 + (size_t)(num_slots_ - 1), nullptr);
  if (init_resources() != Status::SUCCESS) cleanup_resources();     // compress_stream_pipeline then reports ERROR_NOT_INITIALIZED
}

PipelinedBatchManager::~PipelinedBatchManager() { cleanup_resources(); }

Status PipelinedBatchManager::init_resources() {
  for (auto &st : streams_)
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); return Status::ERROR_CUDA_ERROR; }
  // + frame header of the multi-block form; rounded so that the 16-byte result mailbox behind it is aligned
  const size_t out_cap = (manager_->get_max_compressed_size(batch_size_) + 64 + 15) & ~(size_t)15;
  const size_t ws_cap = manager_->get_compress_temp_size(batch_size_);
  for (auto &s : ring_buffer_) {
    s.input_capacity = batch_size_; s.output_capacity = out_cap; s.workspace_capacity = ws_cap;
    bool ok = cudaMalloc(&s.d_input, batch_size_) == cudaSuccess && cudaMalloc(&s.d_output, out_cap) == cudaSuccess &&
              cudaMalloc(&s.d_workspace, ws_cap) == cudaSuccess && cudaMallocHost(&s.h_input, batch_size_) == cudaSuccess &&
              cudaMallocHost(&s.h_output, out_cap + 16) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.event_uploaded, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s.event_compressed, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s.event_downloaded, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { (void)cudaGetLastError(); return Status::ERROR_OUT_OF_MEMORY; }
  }
  return Status::SUCCESS;
}

void PipelinedBatchManager::cleanup_resources() {
  for (auto &s : ring_buffer_) {
    if (s.d_input) cudaFree(s.d_input);
    if (s.d_output) cudaFree(s.d_output);
    if (s.d_workspace) cudaFree(s.d_workspace);
    if (s.h_input) cudaFreeHost(s.h_input);
    if (s.h_output) cudaFreeHost(s.h_output);
    if (s.event_uploaded) cudaEventDestroy(s.event_uploaded);
    if (s.event_compressed) cudaEventDestroy(s.event_compressed);
    if (s.event_downloaded) cudaEventDestroy(s.event_downloaded);
    s = RingBufferSlot();
  }
  for (auto &st : streams_) { if (st) cudaStreamDestroy(st); st = nullptr; }
  (void)cudaGetLastError();
}

Status PipelinedBatchManager::compress_stream_pipeline(std::function<bool(void *, size_t, size_t *)> input_callback,
                                                       std::function<void(const void *, size_t)> output_callback) {
  if (!input_callback || !output_callback) return Status::ERROR_INVALID_PARAMETER;
  if (!manager_ || streams_.size() < 3 || !streams_[0] || ring_buffer_.empty() || !ring_buffer_[0].h_input) return Status::ERROR_NOT_INITIALIZED;
  ZstdBatchManager *codec = batch_of(manager_.get());
  cudaStream_t up = streams_[0], down = streams_[2];
  const size_t S = ring_buffer_.size();
  auto run_of = [&](size_t slot) { return slot == 0 || 2 + slot >= streams_.size() ? streams_[1] : streams_[2 + slot]; };
  // slot states by batch number: batches [drained, fetched) have their D2H in flight, [fetched, issued) are compressing
  size_t issued = 0, fetched = 0, drained = 0;
  Status result = Status::SUCCESS;

  // wait for the size of the oldest compressing batch and start its download
  auto fetch_one = [&]() -> Status {
    RingBufferSlot &s = ring_buffer_[fetched % S];
    if (cudaEventSynchronize(s.event_compressed) != cudaSuccess) return Status::ERROR_CUDA_ERROR;
    const unsigned long long *mb = mailbox(s);
    ++fetched;
    if (mb[1] != 0) { s.current_output_size = 0; return static_cast<Status>((u32)mb[1]); }
    s.current_output_size = (size_t)mb[0];
    if (cudaMemcpyAsync(s.h_output, s.d_output, s.current_output_size, cudaMemcpyDeviceToHost, down) != cudaSuccess) return Status::ERROR_CUDA_ERROR;
    if (cudaEventRecord(s.event_downloaded, down) != cudaSuccess) return Status::ERROR_CUDA_ERROR;
    return Status::SUCCESS;
  };
  // wait for the oldest download and hand the frame to the caller
  auto drain_one = [&]() -> Status {
    RingBufferSlot &s = ring_buffer_[drained % S];
    ++drained;
    if (s.current_output_size == 0) return Status::SUCCESS;        // failed batch: already reported by fetch_one
    if (cudaEventSynchronize(s.event_downloaded) != cudaSuccess) return Status::ERROR_CUDA_ERROR;
    output_callback(s.h_output, s.current_output_size);
    return Status::SUCCESS;
  };

  bool more = true;
  while (more && result == Status::SUCCESS) {
    // the slot of batch `issued` was last used by batch issued - S: it must be fully drained first
    while (result == Status::SUCCESS && issued >= S && drained + S <= issued) {
      if (fetched == drained) result = fetch_one();
      if (result == Status::SUCCESS) result = drain_one();
    }
    if (result != Status::SUCCESS) break;
    RingBufferSlot &s = ring_buffer_[issued % S];
    cudaStream_t run = run_of(issued % S);
    size_t len = 0;
    more = input_callback(s.h_input, s.input_capacity, &len);
    if (len > s.input_capacity) { result = Status::ERROR_INVALID_PARAMETER; break; }
    if (len == 0) continue;                                          // nothing in this batch (end of input, usually)
    s.current_input_size = len; s.current_output_size = 0;
    cudaError_t e = cudaMemcpyAsync(s.d_input, s.h_input, len, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaEventRecord(s.event_uploaded, up);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(run, s.event_uploaded, 0);
    if (e != cudaSuccess) { (void)cudaGetLastError(); result = Status::ERROR_CUDA_ERROR; break; }
    result = codec->compress_async_no_sync(s.d_input, len, s.d_output, s.output_capacity, mailbox(s), s.d_workspace, s.workspace_capacity, run);
    if (result != Status::SUCCESS) break;
    if (cudaEventRecord(s.event_compressed, run) != cudaSuccess) { result = Status::ERROR_CUDA_ERROR; break; }
    ++issued;
    // with this batch on its way, settle older ones: whatever has finished compressing starts its download now, and the
    // host blocks only to keep at most S - 1 batches compressing (one slot is always being filled)
    while (result == Status::SUCCESS && fetched < issued &&
           (fetched + (S - 1) < issued || cudaEventQuery(ring_buffer_[fetched % S].event_compressed) == cudaSuccess))
      result = fetch_one();
    // frames whose download has finished go out as soon as they are there (keeps the caller's sink busy)
    while (result == Status::SUCCESS && drained < fetched && cudaEventQuery(ring_buffer_[drained % S].event_downloaded) == cudaSuccess)
      result = drain_one();
    (void)cudaGetLastError();                                        // cudaErrorNotReady from the query is not an error
  }
  // tail: everything still in flight, in order
  while (result == Status::SUCCESS && fetched < issued) result = fetch_one();
  while (result == Status::SUCCESS && drained < fetched) result = drain_one();
  if (result != Status::SUCCESS) {
    // leave no work behind that still touches the slots
    for (auto st : streams_) cudaStreamSynchronize(st);
    (void)cudaGetLastError();
  }
  return result;
}

} // namespace cuda_zstd

// ---- C ABI ----
struct cuda_zstd_pipeline { cuda_zstd::PipelinedBatchManager *m; };

extern "C" {
cuda_zstd_pipeline_t *cuda_zstd_pipeline_create(int level, int enable_checksum, size_t batch_size_bytes, int num_slots) {
  if (batch_size_bytes == 0 || num_slots < 2) return nullptr;
  cuda_zstd::CompressionConfig c = cuda_zstd::CompressionConfig::from_level(level);
  c.checksum = enable_checksum ? cuda_zstd::ChecksumPolicy::COMPUTE_AND_VERIFY : cuda_zstd::ChecksumPolicy::NO_COMPUTE_NO_VERIFY;
  auto *p = new (std::nothrow) cuda_zstd_pipeline{nullptr};
  if (!p) return nullptr;
  try { p->m = new cuda_zstd::PipelinedBatchManager(c, batch_size_bytes, num_slots); } catch (...) { delete p; return nullptr; }
  return p;
}
void cuda_zstd_pipeline_destroy(cuda_zstd_pipeline_t *p) { if (p) { delete p->m; delete p; } }
int cuda_zstd_pipeline_compress(cuda_zstd_pipeline_t *p, cuda_zstd_pipeline_input_fn in_fn, cuda_zstd_pipeline_output_fn out_fn, void *user) {
  if (!p || !p->m || !in_fn || !out_fn) return static_cast<int>(cuda_zstd::Status::ERROR_INVALID_PARAMETER);
  try {
    return static_cast<int>(p->m->compress_stream_pipeline(
        [&](void *buf, size_t cap, size_t *len) { return in_fn(user, buf, cap, len) != 0; },
        [&](const void *buf, size_t n) { out_fn(user, buf, n); }));
  } catch (...) { return static_cast<int>(cuda_zstd::Status::ERROR_GENERIC); }
}
}
