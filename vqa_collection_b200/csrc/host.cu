// Host-buffer entry point of the whole path (the e2e leg): wire-format HOST buffers in, answers out.
//
// Reference boundary: the `.to(self.device)` copies inside the reference's forwards
// (encoder.py:153-156, predictor.py:82-83, encoder.py:265) followed by Wrapper.forward_vqa
// (wrapper.py:113-118).  The reference ships float32 features (dataset.py:96-104: 295 KB per
// question); at 54 GB/s of PCIe that alone is 5.5 ms per 1024 questions, 13x the forward itself.
// This path therefore packs the features to the resident bf16 format with the HOST cores
// (host_pack.cpp, bit-identical to the device cast) chunk by chunk into pinned staging slots and
// overlaps packing chunk i+1 with the DMA of chunk i, so PCIe carries 2 bytes per feature.
#include <thread>

#include "common.cuh"

extern "C" {
void* vqa_packpool_create(int threads);
void vqa_packpool_destroy(void* p);
int vqa_packpool_threads(void* p);
void vqa_packpool_run(void* p, const float* src, uint16_t* dst, size_t n);
}

namespace vqa {
int cast_f32_to_bf16(const float*, void*, size_t, cudaStream_t);
int split_f32(const float*, void*, void*, size_t, cudaStream_t);
}

using namespace vqa;

namespace {
constexpr int NS = 4;      // pinned staging slots (host pack path)
}

struct vqa_host_ctx {
  void* pool = nullptr;
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_main = nullptr, ev_copies = nullptr;
  void* pinned[NS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t slot_done[NS] = {nullptr, nullptr, nullptr, nullptr};
  bool slot_used[NS] = {false, false, false, false};
  size_t slot_bytes = 0;
  void* d_stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_free[2] = {nullptr, nullptr};
  bool stage_used[2] = {false, false};
  size_t stage_bytes = 0;
  void* d_img = nullptr; size_t d_img_bytes = 0;
  void* d_small = nullptr; size_t d_small_bytes = 0;       // tokens | labels or bbox | label out
  void* h_label = nullptr; size_t h_label_bytes = 0;       // pinned
  cudaEvent_t ev_done = nullptr;                            // answers of the submitted batch are in h_label
  int64_t* pending_out = nullptr; int pending_B = 0;        // submit → wait hand-over
};

static int grow_device(void** p, size_t* have, size_t need) {
  if (*have >= need) return VQA_OK;
  if (*p) VQA_CUDA_CHECK(cudaFree(*p));
  *p = nullptr; *have = 0;
  VQA_CUDA_CHECK(cudaMalloc(p, need));
  *have = need;
  return VQA_OK;
}
static int grow_pinned(void** p, size_t* have, size_t need) {
  if (*have >= need) return VQA_OK;
  if (*p) VQA_CUDA_CHECK(cudaFreeHost(*p));
  *p = nullptr; *have = 0;
  VQA_CUDA_CHECK(cudaHostAlloc(p, need, cudaHostAllocDefault));
  *have = need;
  return VQA_OK;
}

extern "C" {

int vqa_host_ctx_create(vqa_host_ctx** out, int pack_threads) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(out, "vqa_host_ctx_create: NULL out");
  vqa_host_ctx* c = new vqa_host_ctx();
  if (pack_threads <= 0) {
    pack_threads = (int)std::thread::hardware_concurrency();
    if (pack_threads < 1) pack_threads = 1;
  }
  c->pool = vqa_packpool_create(pack_threads);
  cudaError_t e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copies, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming);
  for (int i = 0; i < NS && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->slot_done[i], cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->stage_free[i], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    vqa_packpool_destroy(c->pool);
    delete c;
    return fail(VQA_ERR_CUDA, "vqa_host_ctx_create: %s", cudaGetErrorString(e));
  }
  *out = c;
  return VQA_OK;
}

void vqa_host_ctx_destroy(vqa_host_ctx* c) {
  if (!c) return;
  if (c->copy) cudaStreamSynchronize(c->copy);
  for (int i = 0; i < NS; ++i) { if (c->pinned[i]) cudaFreeHost(c->pinned[i]); if (c->slot_done[i]) cudaEventDestroy(c->slot_done[i]); }
  for (int i = 0; i < 2; ++i) { if (c->d_stage[i]) cudaFree(c->d_stage[i]); if (c->stage_free[i]) cudaEventDestroy(c->stage_free[i]); }
  if (c->d_img) cudaFree(c->d_img);
  if (c->d_small) cudaFree(c->d_small);
  if (c->h_label) cudaFreeHost(c->h_label);
  if (c->ev_main) cudaEventDestroy(c->ev_main);
  if (c->ev_copies) cudaEventDestroy(c->ev_copies);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->copy) cudaStreamDestroy(c->copy);
  vqa_packpool_destroy(c->pool);
  delete c;
}

int vqa_host_ctx_threads(vqa_host_ctx* c) { return c ? vqa_packpool_threads(c->pool) : 0; }

int vqa_forward_host_submit(vqa_host_ctx* c, vqa_forward_host_args* ha, void* stream);

int vqa_forward_host_wait(vqa_host_ctx* c) {
  VQA_REQUIRE(c, "vqa_forward_host_wait: NULL context");
  if (!c->pending_out) return VQA_OK;
  VQA_CUDA_CHECK(cudaEventSynchronize(c->ev_done));
  memcpy(c->pending_out, c->h_label, (size_t)c->pending_B * 8);
  c->pending_out = nullptr;
  c->pending_B = 0;
  return VQA_OK;
}

int vqa_forward_host(vqa_host_ctx* c, vqa_forward_host_args* ha, void* stream) {
  if (int rc = vqa_forward_host_submit(c, ha, stream)) return rc;
  return vqa_forward_host_wait(c);
}

int vqa_forward_host_submit(vqa_host_ctx* c, vqa_forward_host_args* ha, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(c && ha, "vqa_forward_host: NULL context / args");
  VQA_REQUIRE(c->pending_out == nullptr, "vqa_forward_host_submit: the previous batch of this context was not waited for");
  vqa_forward_args a = ha->fwd;
  cudaStream_t main_s = (cudaStream_t)stream;
  VQA_REQUIRE(a.B >= 0 && a.K >= 1 && a.V >= 1 && a.T >= 1, "vqa_forward_host: bad dims");
  VQA_REQUIRE(ha->h_img && ha->h_tokens && ha->h_label, "vqa_forward_host: NULL host buffer");
  VQA_REQUIRE(!a.relation || ha->h_labels || ha->h_bbox, "vqa_forward_host: relation path needs h_labels or h_bbox");
  ha->h2d_bytes = ha->d2h_bytes = 0;
  if (a.B == 0) return VQA_OK;
  const size_t es = elem_size(a.dtype);
  const size_t row_elems = (size_t)a.K * a.V;
  const int chunk = ha->chunk_rows > 0 ? ha->chunk_rows : 64;
  const bool wire_bf16 = ha->img_is_bf16 != 0;
  VQA_REQUIRE(!wire_bf16 || a.dtype == VQA_BF16, "vqa_forward_host: img_is_bf16 needs a bf16 engine");
  const bool pack = a.dtype == VQA_BF16 && ha->pack_on_host && !wire_bf16;
  // hybrid staging: every `raw_chunk_period`-th chunk crosses PCIe as f32 and is cast on the device, the others
  // are packed by the host cores — balances host memory bandwidth (pack) against PCIe bandwidth (raw)
  const int period = pack ? ha->raw_chunk_period : 0;
  VQA_REQUIRE(period >= 0 && period != 1, "vqa_forward_host: raw_chunk_period must be 0 or >= 2");
  int rc;
  // ---- device / pinned buffers owned by the context
  if ((rc = grow_device(&c->d_img, &c->d_img_bytes, (size_t)a.B * row_elems * es))) return rc;
  const size_t tok_bytes = align_up((size_t)a.B * a.T * 8, 256);
  const size_t lab_bytes = a.relation ? align_up(ha->h_labels ? (size_t)a.B * a.K * a.K : (size_t)a.B * a.K * 16, 256) : 0;
  const size_t out_lab_bytes = a.relation && !ha->h_labels ? align_up((size_t)a.B * a.K * a.K, 256) : 0;
  const size_t ans_bytes = align_up((size_t)a.B * 8, 256);
  if ((rc = grow_device(&c->d_small, &c->d_small_bytes, tok_bytes + lab_bytes + out_lab_bytes + ans_bytes))) return rc;
  if ((rc = grow_pinned(&c->h_label, &c->h_label_bytes, (size_t)a.B * 8))) return rc;
  char* small = (char*)c->d_small;
  int64_t* d_tokens = (int64_t*)small;
  void* d_lab_in = small + tok_bytes;
  uint8_t* d_lab_out = (uint8_t*)(small + tok_bytes + lab_bytes);
  int64_t* d_label = (int64_t*)(small + tok_bytes + lab_bytes + out_lab_bytes);
  if (pack) {
    const size_t need = (size_t)chunk * row_elems * 2;
    if (c->slot_bytes < need) {
      for (int i = 0; i < NS; ++i) {
        if (c->slot_used[i]) { VQA_CUDA_CHECK(cudaEventSynchronize(c->slot_done[i])); c->slot_used[i] = false; }
        if (c->pinned[i]) { VQA_CUDA_CHECK(cudaFreeHost(c->pinned[i])); c->pinned[i] = nullptr; }
        VQA_CUDA_CHECK(cudaHostAlloc(&c->pinned[i], need, cudaHostAllocDefault));
      }
      c->slot_bytes = need;
    }
  }
  const bool split = a.dtype == VQA_F16X2;             // fp32-class engine: f32 over PCIe, fp16 plane pair made on the device
  if (split || (a.dtype == VQA_BF16 && !wire_bf16 && (!pack || period > 0))) {
    const size_t need = (size_t)chunk * row_elems * 4;
    if (c->stage_bytes < need) {
      for (int i = 0; i < 2; ++i) {
        if (c->d_stage[i]) { VQA_CUDA_CHECK(cudaFree(c->d_stage[i])); c->d_stage[i] = nullptr; }
        VQA_CUDA_CHECK(cudaMalloc(&c->d_stage[i], need));
        c->stage_used[i] = false;
      }
      c->stage_bytes = need;
    }
  }
  // the context's buffers are private and its previous batch has been waited for (ev_done), so the copies
  // need not wait for whatever else is queued on `stream` (e.g. the forward of ANOTHER context's batch:
  // that is what lets batch n+1 be staged while batch n computes)
  // ---- features, chunk by chunk
  int i = 0;
  for (int b0 = 0; b0 < a.B; b0 += chunk, ++i) {
    const int rows = a.B - b0 < chunk ? a.B - b0 : chunk;
    const size_t n = (size_t)rows * row_elems;
    const float* src = ha->h_img + (size_t)b0 * row_elems;
    char* dst = (char*)c->d_img + (size_t)b0 * row_elems * es;
    const bool raw = period > 0 && (i % period == period - 1);
    if (wire_bf16) {                                   // already in the resident format: one DMA per chunk
      VQA_CUDA_CHECK(cudaMemcpyAsync(dst, (const char*)ha->h_img + (size_t)b0 * row_elems * 2, n * 2, cudaMemcpyHostToDevice, c->copy));
      ha->h2d_bytes += n * 2;
    } else if (pack && !raw) {
      const int slot = i % NS;
      if (c->slot_used[slot]) VQA_CUDA_CHECK(cudaEventSynchronize(c->slot_done[slot]));      // its previous DMA has drained
      vqa_packpool_run(c->pool, src, (uint16_t*)c->pinned[slot], n);
      VQA_CUDA_CHECK(cudaMemcpyAsync(dst, c->pinned[slot], n * 2, cudaMemcpyHostToDevice, c->copy));
      VQA_CUDA_CHECK(cudaEventRecord(c->slot_done[slot], c->copy));
      c->slot_used[slot] = true;
      ha->h2d_bytes += n * 2;
    } else if (split) {
      const int sl = i & 1;
      VQA_CUDA_CHECK(cudaMemcpyAsync(c->d_stage[sl], src, n * 4, cudaMemcpyHostToDevice, c->copy));
      char* hi = (char*)c->d_img + (size_t)b0 * row_elems * 2;
      if ((rc = split_f32((const float*)c->d_stage[sl], hi, hi + (size_t)a.B * row_elems * 2, n, c->copy))) return rc;
      ha->h2d_bytes += n * 4;
    } else if (a.dtype == VQA_BF16) {
      // raw chunk: f32 over PCIe, cast on the device.  The cast runs on the COPY stream right behind its DMA (in-order, so
      // the two staging buffers need no events) — on the main stream it would queue behind the forward of the batch that
      // is still computing and stall this batch's DMA stream after two chunks.
      const int sl = (period > 0 ? i / period : i) & 1;
      VQA_CUDA_CHECK(cudaMemcpyAsync(c->d_stage[sl], src, n * 4, cudaMemcpyHostToDevice, c->copy));
      if ((rc = cast_f32_to_bf16((const float*)c->d_stage[sl], dst, n, c->copy))) return rc;
      ha->h2d_bytes += n * 4;
    } else {
      VQA_CUDA_CHECK(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyHostToDevice, c->copy));
      ha->h2d_bytes += n * 4;
    }
  }
  // ---- the small inputs
  VQA_CUDA_CHECK(cudaMemcpyAsync(d_tokens, ha->h_tokens, (size_t)a.B * a.T * 8, cudaMemcpyHostToDevice, c->copy));
  ha->h2d_bytes += (size_t)a.B * a.T * 8;
  a.d_labels = nullptr; a.d_bbox = nullptr;
  if (a.relation) {
    if (ha->h_labels) {
      VQA_CUDA_CHECK(cudaMemcpyAsync(d_lab_in, ha->h_labels, (size_t)a.B * a.K * a.K, cudaMemcpyHostToDevice, c->copy));
      ha->h2d_bytes += (size_t)a.B * a.K * a.K;
      a.d_labels = (const uint8_t*)d_lab_in;
    } else {
      VQA_CUDA_CHECK(cudaMemcpyAsync(d_lab_in, ha->h_bbox, (size_t)a.B * a.K * 16, cudaMemcpyHostToDevice, c->copy));
      ha->h2d_bytes += (size_t)a.B * a.K * 16;
      a.d_bbox = (const float*)d_lab_in;
      if (!a.d_labels_out) a.d_labels_out = d_lab_out;
    }
  }
  VQA_CUDA_CHECK(cudaEventRecord(c->ev_copies, c->copy));
  VQA_CUDA_CHECK(cudaStreamWaitEvent(main_s, c->ev_copies, 0));
  // ---- forward on the resident batch, answers back
  a.d_img = c->d_img; a.d_tokens = d_tokens; a.d_label = d_label;
  if ((rc = vqa_forward(&a, stream))) return rc;
  VQA_CUDA_CHECK(cudaMemcpyAsync(c->h_label, d_label, (size_t)a.B * 8, cudaMemcpyDeviceToHost, main_s));
  VQA_CUDA_CHECK(cudaEventRecord(c->ev_done, main_s));
  c->pending_out = ha->h_label;
  c->pending_B = a.B;
  ha->d2h_bytes = (size_t)a.B * 8;
  return VQA_OK;
}

}  // extern "C"
