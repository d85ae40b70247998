// C ABI of the engine (include/text2loc_b200.h): handle, weights, workspace arena and the
// orchestration of the kernels in this directory into encode_cells / encode_text / search.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/text2loc_b200.h"
#include "ops.h"
#include "umma_gemm.cuh"

namespace t2l {

static TmaApi g_tma;
const TmaApi& tma_api() { return g_tma; }
static std::string g_create_error;

struct Weight {
  float* dev = nullptr;
  __half* dev16 = nullptr;  // fp16 copy [rows, ld16] (zero-padded rows, 16-byte pitch) of the weights of the kind::f16 GEMMs
  int rows = 0, cols = 0, ld = 0, ld16 = 0;
};

// bump allocator over one device arena, reset per chunk
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0;
  bool overflow = false;
  template <class T>
  T* get(size_t n) {
    off = (off + 255) & ~size_t(255);
    if (off + n * sizeof(T) > cap) {  // never hand out an out-of-bounds pointer; the caller fails the call
      overflow = true;
      return reinterpret_cast<T*>(base);
    }
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace t2l

using namespace t2l;

struct t2l_engine {
  int device = 0;
  std::string err;
  std::map<std::string, Weight> w;
  bool finalized = false;
  bool fine = false;         // weights are those of the fine-stage model (CrossMatch, d = 128): only t2l_fine_offsets runs
  Arena arena;
  Launches lc;
  SearchDb db;
  SearchWork sw{};
  size_t sw_planes_rows = 0;
  int obj_chunk = 16384;     // objects per encode chunk (cell-aligned); 2048 -> 4096 -> 8192 -> 16384 is +9 % / +5 % / +4 % cells/s (fuller
                             // grids for the small kernels), ~6.5 GB of workspace reserved (T2L_OBJ_CHUNK to change)
  bool dist_fma = false;     // FPS / ball-query distances with FMA contraction (T2L_DIST_FMA=1; oracle: pyg_ops.DIST_FMA)
  int tok_chunk = 75776;     // tokens per text chunk (sentence-aligned): 296 row tiles of 256 = whole waves of 74 CTA pairs for all four
                             // token GEMMs; 32768 -> 75776 is ~5 % on the text head (fewer launch tails), larger gains nothing
                             // (scripts/tok_chunk_sweep.py); T2L_TOK_CHUNK to change
  bool text_f16 = true;      // token layer on fp16 operands (same 11-bit significand as tf32, twice the MMA rate, half the
                             // operand bytes); T2L_TEXT_TF32=1 selects the tf32 path for A/B checks
  bool text_stream16 = true; // token layer's residual stream (x + attn, LayerNorm1, x1 + ffn) carried as fp16 rows between the GEMMs and
                             // the LayerNorms; T2L_TEXT_STREAM32=1 keeps it in fp32 (round 1's layout)
  int sa_bisect = 0;         // t2l_debug_sa_bisect: timing bisect of the fused set-abstraction kernel (results invalid when != 0)
  bool text_reg_epilogue = false;  // T2L_TEXT_REG_EPI=1: register-staged residual epilogue instead of the TMA one (A/B)
  float* pooled = nullptr;   // [pooled_cap, 1024] max-over-tokens sentence features between the two text stages
  size_t pooled_cap = 0;
  // pinned staging ring for the per-chunk index arrays of encode_cells: a pageable cudaMemcpyAsync larger than 64 KB
  // serialises the host with the stream, so the host could not queue chunk k+1 while chunk k ran (encode times then
  // varied 60 -> 95 ms between identical runs)
  struct HostStage { int32_t* ptr = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; bool used = false; } stage[4];
  unsigned stage_next = 0;
  // The arena, `pooled` and the search work buffers are shared by every call on this engine.  Calls on ONE stream are
  // ordered by the stream; a call on a different stream first waits for the event the previous call recorded.
  cudaEvent_t order_ev = nullptr;
  cudaStream_t last_stream = nullptr;
  bool order_recorded = false;
  // Which candidate pass the search runs first.  The single fp16 pass is a third of the tensor work but its error bound is
  // ~6x wider; when most queries of a call fail its proof (tightly clustered databases) the three-pass bf16 product is the
  // cheaper FIRST pass.  The count of the previous call is read back asynchronously (pinned word + event, never waited
  // for) and steers the next call; the returned top-k is identical either way.  T2L_SEARCH_FIRST=fp16|bf16x3 pins it.
  int search_first = 0;        // 0 adaptive, 1 always fp16, 2 always bf16x3
  bool bf16_first_now = false;
  int bf16_first_calls = 0;
  int32_t* fail_host = nullptr;
  float* fine_buf = nullptr;   // object encodings | hint encodings between the stages of t2l_fine_offsets
  size_t fine_cap = 0;
  int64_t* acc_idx = nullptr;  // per-chunk lists of t2l_search_topk_accumulate
  double* acc_score = nullptr;
  int acc_cap = 0;
  cudaEvent_t fail_ev = nullptr;
  bool fail_pending = false;
  int fail_nq = 0;
};

namespace {
// Every entry point runs on the engine's device and restores the caller's current device on return.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess; else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
// Cross-stream ordering of the shared workspace (see t2l_engine::order_ev).
struct CallScope {
  t2l_engine* e;
  cudaStream_t st;
  bool ok = true;
  bool capturing = false;  // inside a CUDA graph capture the stream order IS the graph order: no cross-stream events are recorded
  CallScope(t2l_engine* e_, void* stream) : e(e_), st(static_cast<cudaStream_t>(stream)) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess) capturing = cap != cudaStreamCaptureStatusNone;
    if (!capturing && e->order_recorded && st != e->last_stream) ok = cudaStreamWaitEvent(st, e->order_ev, 0) == cudaSuccess;
  }
  ~CallScope() {
    if (capturing) return;
    if (cudaEventRecord(e->order_ev, st) == cudaSuccess) { e->last_stream = st; e->order_recorded = true; }
  }
};
}  // namespace

static int fail(t2l_engine* e, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return 1;
}

#define ENTER(e)                     \
  DeviceGuard _guard((e)->device);   \
  if (!_guard.ok) return fail(e, "cudaSetDevice(%d) failed", (e)->device)
#define ENTER_STREAM(e, stream)      \
  ENTER(e);                          \
  CallScope _scope(e, stream);       \
  if (!_scope.ok) return fail(e, "cudaStreamWaitEvent failed")

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _c = (call);                                                                       \
    if (_c != cudaSuccess) return fail(e, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_c), __FILE__, __LINE__); \
  } while (0)

static int ensure_arena(t2l_engine* e, size_t bytes) {
  e->arena.overflow = false;
  if (e->arena.cap >= bytes) { e->arena.off = 0; return 0; }
  if (e->arena.base) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->arena.base)); e->arena.base = nullptr; e->arena.cap = 0; }
  bytes += bytes / 8;
  CU(cudaMalloc(&e->arena.base, bytes));
  e->arena.cap = bytes;
  e->arena.off = 0;
  return 0;
}

extern "C" int t2l_version(void) { return 1; }

extern "C" const char* t2l_last_error(const t2l_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" int64_t t2l_launch_count(const t2l_engine* e) { return e ? e->lc.n : 0; }

extern "C" int t2l_create(int device, t2l_engine** out) {
  t2l_engine* e = nullptr;
  if (!out) return fail(e, "t2l_create: out is NULL");
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(e, "t2l_create: no CUDA device visible; this engine has no CPU path");
  if (device < 0 || device >= n_dev) return fail(e, "t2l_create: device %d out of range (%d visible)", device, n_dev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(e, "t2l_create: cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(e, "t2l_create: device %d is sm_%d%d; the kernels are sm_100a (Blackwell B200) only", device, prop.major, prop.minor);
  if (!g_tma.encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(e, "t2l_create: cuTensorMapEncodeTiled not available in this driver");
    g_tma.encode = reinterpret_cast<TmaApi::EncodeTiled>(fn);
  }
  g_tma.num_sms = prop.multiProcessorCount;
  e = new t2l_engine();
  e->device = device;
  if (const char* v = getenv("T2L_TEXT_TF32")) e->text_f16 = !(v[0] == '1');
  if (const char* v = getenv("T2L_TEXT_STREAM32")) e->text_stream16 = !(v[0] == '1');
  if (const char* v = getenv("T2L_TEXT_REG_EPI")) e->text_reg_epilogue = v[0] == '1';
  if (const char* v = getenv("T2L_DIST_FMA")) e->dist_fma = v[0] == '1';
  if (const char* v = getenv("T2L_TOK_CHUNK")) { const int n = atoi(v); if (n >= 256 && n <= (1 << 20)) e->tok_chunk = n; }
  if (const char* v = getenv("T2L_OBJ_CHUNK")) { const int n = atoi(v); if (n >= 64 && n <= 65536) e->obj_chunk = n; }
  if (const char* v = getenv("T2L_SEARCH_FIRST")) e->search_first = v[0] == 'f' ? 1 : (v[0] == 'b' ? 2 : 0);
  if (cudaMalloc(&e->db.max_norm, sizeof(float)) != cudaSuccess || cudaMalloc(&e->db.scale, sizeof(float)) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&e->fail_host), sizeof(int32_t)) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->fail_ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->order_ev, cudaEventDisableTiming) != cudaSuccess) {
    delete e;
    return fail(nullptr, "t2l_create: cudaMalloc / cudaEventCreate failed");
  }
  *out = e;
  return 0;
}

static void free_search_work(t2l_engine* e) {
  cudaFree(e->sw.q_planes); cudaFree(e->sw.q16); cudaFree(e->sw.q_scale); cudaFree(e->sw.q_norm); cudaFree(e->sw.cand_score); cudaFree(e->sw.cand_idx);
  cudaFree(e->sw.cand_thr); cudaFree(e->sw.flags); cudaFree(e->sw.n_fail); cudaFree(e->sw.fail_ids); cudaFree(e->sw.fail_thr);
  cudaFree(e->sw.q2_planes); cudaFree(e->sw.cand2_idx); cudaFree(e->sw.cand2_cnt);
  e->sw = SearchWork{};
}

extern "C" void t2l_destroy(t2l_engine* e) {
  if (!e) return;
  DeviceGuard guard(e->device);
  cudaDeviceSynchronize();
  for (auto& kv : e->w) { cudaFree(kv.second.dev); cudaFree(kv.second.dev16); }
  cudaFree(e->arena.base);
  cudaFree(e->pooled);
  for (auto& hs : e->stage) { if (hs.ptr) cudaFreeHost(hs.ptr); if (hs.ev) cudaEventDestroy(hs.ev); }
  if (e->order_ev) cudaEventDestroy(e->order_ev);
  if (e->fail_ev) cudaEventDestroy(e->fail_ev);
  if (e->fail_host) cudaFreeHost(e->fail_host);
  cudaFree(e->db.plane16);
  cudaFree(e->db.scale);
  cudaFree(e->acc_idx);
  cudaFree(e->acc_score);
  cudaFree(e->fine_buf);
  cudaFree(e->db.planes);
  cudaFree(e->db.max_norm);
  free_search_work(e);
  delete e;
}

// ---------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------
static bool is_tf32_operand(const std::string& n) {
  // weights consumed by the tcgen05 tf32 GEMMs: pre-rounded once (round-to-nearest) so the tensor
  // core's own truncation of the B operand is exact
  static const char* pre[] = {"sa1.w2", "sa2.w1x", "sa2.w1q", "sa2.w2", "sa3.w1x", "sa3.w1q", "sa3.w2", "ga.w1", "ga.w2", "lin1.w", "lin2.w",
                              "mlp_pointnet.w", "merge.w", "txt_intra.in_w", "txt_intra.out_w", "txt_intra.l1_w", "txt_intra.l2_w"};
  for (const char* p : pre) if (n == p) return true;
  return false;
}

static bool is_f16_operand(const std::string& n) {
  return n == "txt_intra.in_w" || n == "txt_intra.out_w" || n == "txt_intra.l1_w" || n == "txt_intra.l2_w" || n == "sa1.w2" || n == "sa2.w2" ||
         n == "sa3.w2" || n == "ga.w1" || n == "ga.w2";
}

static bool is_split3_operand(const std::string& n) {
  // weights of the fp32-accurate tensor-core layers: stored as [hi | lo] tf32 planes
  static const char* pre[] = {"obj_attn0", "obj_attn1", "txt_inter", "cross_objects0", "cross_objects1", "cross_hints0", "cross_hints1"};
  static const char* parts[] = {".in_w", ".out_w", ".l1_w", ".l2_w", ".ca_q_w", ".ca_kv_w", ".ca_out_w"};
  for (const char* p : pre) for (const char* q : parts) if (n == std::string(p) + q) return true;
  return n == "txt_mlp.w";
}

static float host_round_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & ~0x1fffu;  // rna: add half ulp of the 10-bit mantissa, truncate
  memcpy(&x, &u, 4);
  return x;
}

extern "C" int t2l_set_weight(t2l_engine* e, const char* name, const float* data, int rows, int cols) {
  if (!e || !name || !data || rows <= 0 || cols <= 0) return fail(e, "t2l_set_weight: bad argument");
  ENTER(e);
  Weight& w = e->w[name];
  if (w.dev) { CU(cudaFree(w.dev)); w.dev = nullptr; }
  if (w.dev16) { CU(cudaFree(w.dev16)); w.dev16 = nullptr; }
  const std::string nm(name);
  if (nm.size() > 4 && nm.compare(nm.size() - 4, 4, ".w1p") == 0) {
    // sa_obj2.cu forms v = W1p . (pos_i - o) in fp16 next to Qx (|Qx| <= 32752): with in-cell position differences of at most
    // 4 units (cell-normalised coordinates lie in [0, 1], NormalizeScale'd ones in (-1, 1)) |v| must stay below 32752 as well
    for (int r = 0; r < rows; ++r) {
      double l1 = 0;
      for (int c = 0; c < cols; ++c) l1 += fabs(static_cast<double>(data[static_cast<size_t>(r) * cols + c]));
      if (!(4.0 * l1 <= 32752.0)) return fail(e, "t2l_set_weight: '%s' row %d is too large for the fp16 set-abstraction path", name, r);
    }
  }
  const bool split3 = is_split3_operand(name);
  if (split3 && (cols % 32)) return fail(e, "t2l_set_weight: '%s' needs cols %% 32 == 0", name);
  w.rows = rows; w.cols = cols; w.ld = split3 ? 2 * cols : ((cols + 3) & ~3);
  std::vector<float> host(static_cast<size_t>(rows) * w.ld, 0.f);
  const bool rnd = is_tf32_operand(name);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const float v = data[static_cast<size_t>(r) * cols + c];
      if (split3) {
        const float hi = host_round_tf32(v);
        host[static_cast<size_t>(r) * w.ld + c] = hi;
        host[static_cast<size_t>(r) * w.ld + cols + c] = host_round_tf32(v - hi);
      } else {
        host[static_cast<size_t>(r) * w.ld + c] = rnd ? host_round_tf32(v) : v;
      }
    }
  CU(cudaMalloc(&w.dev, host.size() * sizeof(float)));
  CU(cudaMemcpy(w.dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (is_f16_operand(name)) {
    w.ld16 = (cols + 7) & ~7;
    std::vector<__half> h16(static_cast<size_t>(rows) * w.ld16, __float2half_rn(0.f));
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) {
        const float v = data[static_cast<size_t>(r) * cols + c];
        if (!(fabsf(v) <= 65504.f)) return fail(e, "t2l_set_weight: '%s' has a value outside the fp16 range", name);
        h16[static_cast<size_t>(r) * w.ld16 + c] = __float2half_rn(v);
      }
    CU(cudaMalloc(&w.dev16, h16.size() * sizeof(__half)));
    CU(cudaMemcpy(w.dev16, h16.data(), h16.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  e->finalized = false;
  return 0;
}

// Expected (rows, cols) of every weight, as the reference's constructors fix them (models/pointcloud/pointnet2.py:57-63,
// models/object_encoder.py:33-64, models/cell_retrieval.py:35, models/language_encoder.py:98-103, models/cross_matcher.py:55-78)
// for embed dim d (256 coarse / 128 fine).  A checkpoint with other dimensions is rejected here -- the reference's
// load_state_dict raises on a size mismatch even with strict=False -- instead of being read out of bounds by kernels whose
// leading dimensions are compile-time constants.
struct ShapeSpec { std::string name; int rows, cols; };

static void attn_shapes(std::vector<ShapeSpec>& v, const std::string& p, int d, int ffn) {
  v.push_back({p + ".in_w", 3 * d, d}); v.push_back({p + ".in_b", 1, 3 * d});
  v.push_back({p + ".out_w", d, d});    v.push_back({p + ".out_b", 1, d});
  v.push_back({p + ".l1_w", ffn, d});   v.push_back({p + ".l1_b", 1, ffn});
  v.push_back({p + ".l2_w", d, ffn});   v.push_back({p + ".l2_b", 1, d});
  v.push_back({p + ".n1_w", 1, d}); v.push_back({p + ".n1_b", 1, d}); v.push_back({p + ".n2_w", 1, d}); v.push_back({p + ".n2_b", 1, d});
}

static std::vector<ShapeSpec> expected_shapes(bool fine) {
  const int d = fine ? T2L_FINE_DIM : T2L_EMBED_DIM;
  std::vector<ShapeSpec> v;
  const int sa[3][3] = {{3, 32, 64}, {64, 128, 128}, {128, 256, 256}};  // in, C1, C2
  for (int i = 0; i < 3; ++i) {
    const std::string p = "sa" + std::to_string(i + 1);
    v.push_back({p + ".w1x", sa[i][1], sa[i][0]}); v.push_back({p + ".w1p", sa[i][1], 3}); v.push_back({p + ".b1", 1, sa[i][1]});
    v.push_back({p + ".w2", sa[i][2], sa[i][1]});  v.push_back({p + ".b2", 1, sa[i][2]});
    if (i > 0) v.push_back({p + ".w1q", sa[i][1], sa[i][0] + 6});  // [W1x | W1p | W1p]: the per-point Linear incl. its position part
  }
  v.push_back({"ga.w1", 512, 259}); v.push_back({"ga.b1", 1, 512}); v.push_back({"ga.w2", 1024, 512}); v.push_back({"ga.b2", 1, 1024});
  v.push_back({"lin1.w", 512, 1024}); v.push_back({"lin1.b", 1, 512}); v.push_back({"lin2.w", 256, 512}); v.push_back({"lin2.b", 1, 256});
  v.push_back({"mlp_pointnet.w", d, 256}); v.push_back({"mlp_pointnet.b", 1, d});
  const char* side[3] = {"color", "pos", "num"};
  for (int i = 0; i < 3; ++i) {
    const std::string p = side[i];
    v.push_back({p + ".w1", 64, i == 2 ? 1 : 3}); v.push_back({p + ".b1", 1, 64}); v.push_back({p + ".w2", d, 64}); v.push_back({p + ".b2", 1, d});
  }
  v.push_back({"merge.w", d, 4 * d}); v.push_back({"merge.b", 1, d});
  v.push_back({"txt_mlp.w", d, T2L_T5_DIM}); v.push_back({"txt_mlp.b", 1, d});
  attn_shapes(v, "txt_intra", T2L_T5_DIM, 4 * T2L_T5_DIM);
  if (!fine) {
    attn_shapes(v, "obj_attn0", d, 2 * d);
    attn_shapes(v, "obj_attn1", d, 2 * d);
    attn_shapes(v, "txt_inter", d, 4 * d);
  } else {
    for (const char* side_name : {"cross_objects", "cross_hints"})
      for (int i = 0; i < 2; ++i) {
        const std::string p = std::string(side_name) + std::to_string(i);
        attn_shapes(v, p, d, 4 * d);  // self-attention block + FFN + norm1 / norm2 (norm2 follows the cross-attention)
        v.push_back({p + ".ca_q_w", d, d});       v.push_back({p + ".ca_q_b", 1, d});       // multihead_attn.in_proj rows [0, d)
        v.push_back({p + ".ca_kv_w", 2 * d, d});  v.push_back({p + ".ca_kv_b", 1, 2 * d});  // rows [d, 3d): keys | values of the memory
        v.push_back({p + ".ca_out_w", d, d});     v.push_back({p + ".ca_out_b", 1, d});
        v.push_back({p + ".n3_w", 1, d}); v.push_back({p + ".n3_b", 1, d});
      }
    v.push_back({"offs.w1", d / 2, d}); v.push_back({"offs.b1", 1, d / 2}); v.push_back({"offs.w2", 2, d / 2}); v.push_back({"offs.b2", 1, 2});
  }
  return v;
}

extern "C" int t2l_finalize_weights(t2l_engine* e) {
  if (!e) return 1;
  const bool fine = e->w.count("offs.w1") != 0;  // the fine-stage model (CrossMatch) carries the offset MLP
  for (const ShapeSpec& sp : expected_shapes(fine)) {
    auto it = e->w.find(sp.name);
    if (it == e->w.end()) return fail(e, "t2l_finalize_weights: missing weight '%s' (%s model)", sp.name.c_str(), fine ? "fine" : "coarse");
    if (it->second.rows != sp.rows || it->second.cols != sp.cols)
      return fail(e, "t2l_finalize_weights: size mismatch for '%s': got [%d, %d], the engine is built for [%d, %d]", sp.name.c_str(),
                  it->second.rows, it->second.cols, sp.rows, sp.cols);
  }
  ENTER(e);
  CU(cudaDeviceSynchronize());
  e->fine = fine;
  e->finalized = true;
  return 0;
}

static const Weight& W(t2l_engine* e, const std::string& n) { return e->w.at(n); }

// y = act(x W^T + b) helper
static cudaError_t lin(t2l_engine* e, bool umma, const float* A, long lda, int M, const std::string& wname, const std::string& bname,
                       float* C, long ldc, int act, cudaStream_t st, const float* residual = nullptr, long ldr = 0, int round_out = 0,
                       int segmax = 0, const float* side = nullptr, long lds = 0, int out_half = 0, float half_max = 65504.f) {
  const Weight& w = W(e, wname);
  Linear l;
  l.A = A; l.lda = lda; l.W = w.dev; l.ldw = w.ld; l.bias = bname.empty() ? nullptr : W(e, bname).dev;
  l.C = C; l.ldc = ldc; l.M = M; l.N = w.rows; l.K = w.cols; l.act = act;
  l.residual = residual; l.ldr = ldr; l.round_out = round_out; l.segmax = segmax; l.side = side; l.lds = lds; l.out_half = out_half;
  l.half_max = half_max;
  return umma ? linear_umma(l, st, &e->lc) : linear_simt(l, st, &e->lc);
}

// fp32-accurate layer on the tensor cores: split A into [hi | lo] tf32 planes, three-pass GEMM
// against the weight's planes (hi*hi + lo*hi + hi*lo).  Error ~2^-21 per product, i.e. fp32 level.
static cudaError_t lin3(t2l_engine* e, const float* A, long lda, int M, const std::string& wname, const std::string& bname, float* C,
                        long ldc, int act, cudaStream_t st, const float* residual = nullptr, long ldr = 0, const float* A_planes = nullptr,
                        int split_out = 0) {
  // A_planes: the producer of A already wrote its [hi | lo] planes ([M, 2K]; LayerNorm, the attention cores and the split_out
  // epilogue do): no split kernel.  split_out: C receives the planes of the result ([M, ldc >= 2N]) for the next such layer.
  const Weight& w = W(e, wname);
  if (!A_planes) {
    float* planes = e->arena.get<float>(static_cast<size_t>(M) * 2 * w.cols);
    cudaError_t err = split_tf32_planes(A, lda, planes, M, w.cols, st, &e->lc);
    if (err != cudaSuccess) return err;
    A_planes = planes;
  }
  Linear l;
  l.A = A_planes; l.lda = 2L * w.cols; l.W = w.dev; l.ldw = w.ld; l.bias = bname.empty() ? nullptr : W(e, bname).dev;
  l.C = C; l.ldc = ldc; l.M = M; l.N = w.rows; l.K = w.cols; l.act = act; l.residual = residual; l.ldr = ldr; l.passes = 3;
  l.split_out = split_out;
  return linear_umma(l, st, &e->lc);
}

// y = act(x W^T + b) with fp16 operands (A and the weight's fp16 copy), fp32 accumulate; C is fp32 or fp16 (out_half).
static cudaError_t lin_h(t2l_engine* e, const __half* A, long lda, int M, const std::string& wname, const std::string& bname, void* C,
                         long ldc, int act, int out_half, cudaStream_t st, const float* residual = nullptr, long ldr = 0, int segmax = 0,
                         int residual_half = 0) {
  const Weight& w = W(e, wname);
  if (!w.dev16) return cudaErrorInvalidValue;
  Linear l;
  l.A = reinterpret_cast<const float*>(A); l.lda = lda; l.W = reinterpret_cast<const float*>(w.dev16); l.ldw = w.ld16;
  l.bias = bname.empty() ? nullptr : W(e, bname).dev;
  l.C = static_cast<float*>(C); l.ldc = ldc; l.M = M; l.N = w.rows; l.K = w.cols; l.act = act; l.residual = residual; l.ldr = ldr;
  l.half_ops = 1; l.out_half = out_half; l.segmax = segmax; l.round_out = segmax ? 1 : 0; l.residual_half = residual_half;
  l.reg_epilogue = e->text_reg_epilogue ? 1 : 0;
  return linear_umma(l, st, &e->lc);
}

// One post-norm nn.TransformerEncoderLayer on packed rows [n_seq * S, d] (sequence-major).
// fast: single-pass tf32 projections (text token layer, where the FLOPs are); otherwise the
// three-pass split product, which keeps fp32 accuracy (the object and sentence layers amplify
// operand rounding the most, DESIGN.md precision table).
// pooled_out != nullptr: the layer's output is only needed max-pooled over each sequence (token layer): norm2 and the max
// are one kernel and Xout is not written.
// x_is_half: X points at __half rows (fp16 T5 states): they are the first GEMM's A operand as they are and the residual of the
// out-projection is read from them (no fp32 copy of the input exists).
static int encoder_layer(t2l_engine* e, const std::string& pfx, bool fast, const float* X, float* Xout, int n_seq, int S, int d, int ffn,
                         cudaStream_t st, float* pooled_out = nullptr, bool x_is_half = false) {
  const int rows = n_seq * S;
  Arena& a = e->arena;
  float* qkv = a.get<float>(static_cast<size_t>(rows) * 3 * d);
  float* att = a.get<float>(static_cast<size_t>(rows) * d);
  float* y = a.get<float>(static_cast<size_t>(rows) * d);
  float* x1 = a.get<float>(static_cast<size_t>(rows) * d);
  float* h = a.get<float>(static_cast<size_t>(rows) * ffn);
  if (fast && e->text_f16 && d == 1024) {
    // fp16 operands, fp32 accumulation and fp32 residual/LayerNorm stream: x -> fp16 once, the attention core and
    // LayerNorm1 emit the fp16 A operands of the next GEMMs, the FFN hidden activations exist only in fp16
    const __half* xh = reinterpret_cast<const __half*>(X);
    __half* atth = reinterpret_cast<__half*>(att);
    __half* x1h = a.get<__half>(static_cast<size_t>(rows) * d);
    __half* hh = reinterpret_cast<__half*>(h);
    if (!x_is_half) {
      __half* xc = a.get<__half>(static_cast<size_t>(rows) * d);
      CU(to_half_rows(X, xc, static_cast<long>(rows) * d, st, &e->lc));
      xh = xc;
    }
    CU(lin_h(e, xh, d, rows, pfx + ".in_w", pfx + ".in_b", qkv, 3L * d, 0, /*out_half=*/1, st));
    CU(mha_tc256(qkv, att, n_seq, S, st, &e->lc, /*round_out=*/2, /*half_in=*/1));
    if (e->text_stream16 && pooled_out) {
      // fp16 residual stream: x + attn, LayerNorm1, x1 + ffn exist only as fp16 rows (each sum is formed in fp32 in the GEMM
      // epilogue and rounded once; LayerNorm statistics in fp32).  Halves the bytes of the two HBM-bound epilogues and of both
      // LayerNorm kernels (4.0e-4 worst row error against 2.6e-4 with the fp32 stream, tolerance 1e-3).
      __half* yh = reinterpret_cast<__half*>(y);
      CU(lin_h(e, atth, d, rows, pfx + ".out_w", pfx + ".out_b", yh, d, 0, 1, st, reinterpret_cast<const float*>(xh), d, 0, 1));
      CU(layer_norm_half_rows(yh, x1h, W(e, pfx + ".n1_w").dev, W(e, pfx + ".n1_b").dev, rows, d, st, &e->lc));
      CU(lin_h(e, x1h, d, rows, pfx + ".l1_w", pfx + ".l1_b", hh, ffn, 1, 1, st));
      CU(lin_h(e, hh, ffn, rows, pfx + ".l2_w", pfx + ".l2_b", yh, d, 0, 1, st, reinterpret_cast<const float*>(x1h), d, 0, 1));
      CU(layer_norm_max_half_rows(yh, pooled_out, W(e, pfx + ".n2_w").dev, W(e, pfx + ".n2_b").dev, n_seq, S, d, st, &e->lc));
      return 0;
    }
    CU(lin_h(e, atth, d, rows, pfx + ".out_w", pfx + ".out_b", y, d, 0, 0, st, X, d, 0, x_is_half ? 1 : 0));
    CU(layer_norm_rows(y, x1, W(e, pfx + ".n1_w").dev, W(e, pfx + ".n1_b").dev, rows, d, st, &e->lc, x1h));
    CU(lin_h(e, x1h, d, rows, pfx + ".l1_w", pfx + ".l1_b", hh, ffn, 1, 1, st));
    CU(lin_h(e, hh, ffn, rows, pfx + ".l2_w", pfx + ".l2_b", y, d, 0, 0, st, x1, d));
  } else if (x_is_half) {
    return fail(e, "internal: fp16 input needs the fp16 token layer");
  } else if (fast) {
    CU(lin(e, true, X, d, rows, pfx + ".in_w", pfx + ".in_b", qkv, 3L * d, 0, st));
    if (d == 1024) CU(mha_tc256(qkv, att, n_seq, S, st, &e->lc, /*round_out=*/1));
    else CU(mha_small(qkv, att, n_seq, S, d, 4, st, &e->lc, /*round_out=*/1));
    CU(lin(e, true, att, d, rows, pfx + ".out_w", pfx + ".out_b", y, d, 0, st, X, d));
    CU(layer_norm_rows(y, x1, W(e, pfx + ".n1_w").dev, W(e, pfx + ".n1_b").dev, rows, d, st, &e->lc));
    CU(lin(e, true, x1, d, rows, pfx + ".l1_w", pfx + ".l1_b", h, ffn, 1, st, nullptr, 0, 1));
    CU(lin(e, true, h, ffn, rows, pfx + ".l2_w", pfx + ".l2_b", y, d, 0, st, x1, d));
  } else {
    // the operand planes of the three-pass GEMMs come straight from their producers (attention core, LayerNorm1, the ReLU
    // epilogue of linear1): one split kernel (the layer's input) instead of four
    float* attp = a.get<float>(static_cast<size_t>(rows) * 2 * d);
    float* x1p = a.get<float>(static_cast<size_t>(rows) * 2 * d);
    float* hp = a.get<float>(static_cast<size_t>(rows) * 2 * ffn);
    CU(lin3(e, X, d, rows, pfx + ".in_w", pfx + ".in_b", qkv, 3L * d, 0, st));
    CU(mha_small(qkv, nullptr, n_seq, S, d, 4, st, &e->lc, 0, attp));
    CU(lin3(e, nullptr, d, rows, pfx + ".out_w", pfx + ".out_b", y, d, 0, st, X, d, attp));
    CU(layer_norm_rows(y, x1, W(e, pfx + ".n1_w").dev, W(e, pfx + ".n1_b").dev, rows, d, st, &e->lc, nullptr, x1p));
    CU(lin3(e, nullptr, d, rows, pfx + ".l1_w", pfx + ".l1_b", hp, 2L * ffn, 1, st, nullptr, 0, x1p, /*split_out=*/1));
    CU(lin3(e, nullptr, ffn, rows, pfx + ".l2_w", pfx + ".l2_b", y, d, 0, st, x1, d, hp));
  }
  if (pooled_out && d == 1024) CU(layer_norm_max_rows(y, pooled_out, W(e, pfx + ".n2_w").dev, W(e, pfx + ".n2_b").dev, n_seq, S, d, st, &e->lc));
  else {
    CU(layer_norm_rows(y, Xout, W(e, pfx + ".n2_w").dev, W(e, pfx + ".n2_b").dev, rows, d, st, &e->lc));
    if (pooled_out) CU(max_over_rows(Xout, pooled_out, n_seq, S, d, st, &e->lc));
  }
  return 0;
}

// One intra-cell layer (cell_retrieval.py:101-103: nn.TransformerEncoderLayer(256, 4 heads, ffn 512), no mask over the 28
// zero-padded slots) on the packed rows of a chunk: `rows` = sum over cells of min(n_b, 28) + [n_b < 28] (the padded slots
// of a cell are identical rows; one representative stands for them, weighted as a key by their count).  Three-pass tf32
// projections like encoder_layer's precise branch.
static int cell_attention_layer(t2l_engine* e, const std::string& pfx, const float* X, const float* Xp, float* Xout, float* Xoutp, int rows,
                                int n_cells, const int32_t* row_ptr_dev, const int32_t* cell_ptr_dev, cudaStream_t st) {
  // Xp / Xoutp: tf32 hi | lo planes of the layer's input (from the previous layer's LayerNorm2 or one split kernel) / output
  constexpr int d = T2L_EMBED_DIM, ffn = 512;
  Arena& a = e->arena;
  float* qkv = a.get<float>(static_cast<size_t>(rows) * 3 * d);
  float* attp = a.get<float>(static_cast<size_t>(rows) * 2 * d);
  float* y = a.get<float>(static_cast<size_t>(rows) * d);
  float* x1 = a.get<float>(static_cast<size_t>(rows) * d);
  float* x1p = a.get<float>(static_cast<size_t>(rows) * 2 * d);
  float* hp = a.get<float>(static_cast<size_t>(rows) * 2 * ffn);
  CU(lin3(e, X, d, rows, pfx + ".in_w", pfx + ".in_b", qkv, 3L * d, 0, st, nullptr, 0, Xp));
  CU(mha_cells64(qkv, nullptr, n_cells, row_ptr_dev, cell_ptr_dev, kObjectSlots, d, 4, st, &e->lc, attp));
  CU(lin3(e, nullptr, d, rows, pfx + ".out_w", pfx + ".out_b", y, d, 0, st, X, d, attp));
  CU(layer_norm_rows(y, x1, W(e, pfx + ".n1_w").dev, W(e, pfx + ".n1_b").dev, rows, d, st, &e->lc, nullptr, x1p));
  CU(lin3(e, nullptr, d, rows, pfx + ".l1_w", pfx + ".l1_b", hp, 2L * ffn, 1, st, nullptr, 0, x1p, /*split_out=*/1));
  CU(lin3(e, nullptr, ffn, rows, pfx + ".l2_w", pfx + ".l2_b", y, d, 0, st, x1, d, hp));
  CU(layer_norm_rows(y, Xout, W(e, pfx + ".n2_w").dev, W(e, pfx + ".n2_b").dev, rows, d, st, &e->lc, nullptr, Xoutp));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// encode_cells
// ---------------------------------------------------------------------------------------------
struct ObjDebug {
  float* features2 = nullptr;
  uint8_t *fps1 = nullptr, *fps2 = nullptr, *fps3 = nullptr, *nbr1 = nullptr, *nbr2 = nullptr, *nbr3 = nullptr, *cnt1 = nullptr,
          *cnt2 = nullptr, *cnt3 = nullptr;
};

static size_t obj_chunk_bytes(size_t n, size_t cells) {
  // upper bound of everything encode_chunk carves from the arena: ~330 KB per object (geometry 10 KB, Qx 96 KB, x1..x3 105 KB,
  // GA 99 KB, tails 14 KB) and, per cell, the two attention layers' buffers (~0.6 MB).  A too-small estimate fails the call
  // (Arena::overflow), it never hands out memory out of bounds.
  return n * size_t(400000) + cells * size_t(28) * 256 * 4 * 32 + (size_t(1) << 20);
}

// obj_emb_out != nullptr: stop after the object encoder and write its normalised rows [n_objects, d] (fine stage)
static int encode_chunk(t2l_engine* e, const float* pts, const float* meta, const int32_t* cell_ptr, int c0, int c1, float* out,
                        const ObjDebug* dbg, cudaStream_t st, float* obj_emb_out = nullptr) {
  const int o0 = cell_ptr[c0], o1 = cell_ptr[c1];
  const int n = o1 - o0, B = c1 - c0;
  if (ensure_arena(e, obj_chunk_bytes(n, B))) return 1;
  Arena& a = e->arena;
  const float* p = pts + static_cast<size_t>(o0) * kPoints * 6;

  // self-loop source of each object (PyG add_self_loops on per-cell indices) + local cell_ptr
  const size_t host_n = 2 * static_cast<size_t>(n) + 2 * (static_cast<size_t>(B) + 1);
  auto& hs = e->stage[e->stage_next++ % 4];
  if (hs.used) CU(cudaEventSynchronize(hs.ev));  // the copy that last read this buffer has finished
  if (hs.cap < host_n) {
    if (hs.ptr) CU(cudaFreeHost(hs.ptr));
    hs.ptr = nullptr;
    CU(cudaMallocHost(reinterpret_cast<void**>(&hs.ptr), (host_n + 1024) * sizeof(int32_t)));
    hs.cap = host_n + 1024;
  }
  if (!hs.ev) CU(cudaEventCreateWithFlags(&hs.ev, cudaEventDisableTiming));
  int32_t* host = hs.ptr;
  for (int c = c0; c < c1; ++c)
    for (int o = cell_ptr[c]; o < cell_ptr[c + 1]; ++o) {
      const int b = o - cell_ptr[c];
      host[o - o0] = (cell_ptr[c] - o0) + b / 2;
      host[n + (o - o0)] = b & 1;
    }
  for (int c = c0; c <= c1; ++c) host[2 * static_cast<size_t>(n) + (c - c0)] = cell_ptr[c] - o0;
  // rows of the intra-cell attention layers: min(n_b, 28) objects + ONE row for all 28 - n_b zero-padded slots (mha_seq64_kernel)
  int32_t* rp_host = host + 2 * static_cast<size_t>(n) + B + 1;
  rp_host[0] = 0;
  for (int c = c0; c < c1; ++c) {
    const int nb = cell_ptr[c + 1] - cell_ptr[c];
    rp_host[c - c0 + 1] = rp_host[c - c0] + (nb < kObjectSlots ? nb + 1 : kObjectSlots);
  }
  const int attn_rows = rp_host[B];
  int32_t* d_host = a.get<int32_t>(host_n);
  CU(cudaMemcpyAsync(d_host, host, host_n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CU(cudaEventRecord(hs.ev, st));
  hs.used = true;
  const int32_t* loop_src = d_host;
  const int32_t* loop_half = d_host + n;
  const int32_t* cell_ptr_dev = d_host + 2 * static_cast<size_t>(n);
  const int32_t* row_ptr_dev = cell_ptr_dev + B + 1;

  Geometry g;
  const size_t N = n;
  g.fps1 = a.get<uint8_t>(N * 128); g.fps2 = a.get<uint8_t>(N * 64); g.fps3 = a.get<uint8_t>(N * 32);
  g.cpos1 = a.get<float>(N * 128 * 3); g.cpos2 = a.get<float>(N * 64 * 3); g.cpos3 = a.get<float>(N * 32 * 3);
  g.nbr1 = a.get<uint8_t>(N * 128 * 32); g.nbr2 = a.get<uint8_t>(N * 64 * 32); g.nbr3 = a.get<uint8_t>(N * 32 * 32);
  g.cnt1 = a.get<uint8_t>(N * 128); g.cnt2 = a.get<uint8_t>(N * 64); g.cnt3 = a.get<uint8_t>(N * 32);
  CU(fps_all_levels(p, n, g, e->dist_fma, st, &e->lc));
  CU(ball_query_all_levels(p, n, g, e->dist_fma, st, &e->lc));

  // Set abstraction x3 (pointnet2.py:25-37).  Per level: Qx = W1x x + b1 + W1p (pos - o) per point (fp16, |.| <= 32752), then
  // the fused gather / second Linear / max kernel (sa_obj2.cu).  Rows of x1 / x2 carry 8 extra columns (tf32 hi | lo of
  // pos - o) so that the next level's per-point Linear is ONE tf32 GEMM against [W1x | W1p | W1p].
  __half* qx1 = a.get<__half>(N * 256 * 32);
  __half* qx23 = a.get<__half>(N * 128 * 128);  // levels 2 and 3 (n*128*128 == n*64*256)
  constexpr int ldx1 = 72, ldx2 = 136;
  float* x1 = a.get<float>(N * 128 * ldx1);
  float* x2 = a.get<float>(N * 64 * ldx2);
  float* x3 = a.get<float>(N * 32 * 256);
  struct Level { const char* name; int C1, C2, P, M; float* x; int ldx; const float* dense; const float* cpos;
                 const uint8_t* nbr; const uint8_t* cnt; __half* Qx; float* xout; int ldo; };
  Level lv[3] = {
      {"sa1", 32, 64, 256, 128, nullptr, 0, nullptr, g.cpos1, g.nbr1, g.cnt1, qx1, x1, ldx1},
      {"sa2", 128, 128, 128, 64, x1, ldx1, g.cpos1, g.cpos2, g.nbr2, g.cnt2, qx23, x2, ldx2},
      {"sa3", 256, 256, 64, 32, x2, ldx2, g.cpos2, g.cpos3, g.nbr3, g.cnt3, qx23, x3, 256},
  };
  for (const Level& L : lv) {
    const std::string nm = L.name;
    if (W(e, nm + ".w1p").ld != 4) return fail(e, "internal: w1p pitch");
    if (!L.x) {  // level 1 reads rgb and xyz straight from pts (K = 3 + 3: exact fp32 SIMT)
      if (W(e, nm + ".w1x").ld != 4) return fail(e, "internal: sa1.w1x pitch");
      CU(sa1_qx16(p, n, W(e, nm + ".w1x").dev, W(e, nm + ".w1p").dev, W(e, nm + ".b1").dev, L.Qx, st, &e->lc));
    } else {
      const int c_in = W(e, nm + ".w1x").cols;
      CU(append_pos_cols(L.dense, n, L.P, L.x, L.ldx, c_in, st, &e->lc));
      CU(lin(e, true, L.x, L.ldx, n * L.P, nm + ".w1q", nm + ".b1", reinterpret_cast<float*>(L.Qx), L.C1, 0, st, nullptr, 0, 0, 0, nullptr, 0,
             /*out_half=*/1, kQxMax));
    }
    SaObj2 so;
    so.Qx16 = L.Qx; so.C1 = L.C1; so.C2 = L.C2; so.cpos = L.cpos; so.nbr = L.nbr; so.cnt = L.cnt; so.loop_src_obj = loop_src;
    so.loop_half = loop_half; so.Wp = W(e, nm + ".w1p").dev; so.W2h = W(e, nm + ".w2").dev16; so.b2 = W(e, nm + ".b2").dev;
    so.out = L.xout; so.ldo = L.ldo; so.n_obj = n; so.P = L.P; so.M = L.M; so.bisect = e->sa_bisect;
    if (!so.W2h) return fail(e, "internal: no fp16 copy of %s.w2", L.name);
    CU(sa_obj2(so, st, &e->lc));
  }

  // GlobalAbstraction: mlp([x | pos]) then max over the object's 32 points (pointnet2.py:45-49), fp16 operands like the
  // set-abstraction layers
  __half* gaA16 = a.get<__half>(N * 32 * 264);
  __half* g1h = a.get<__half>(N * 32 * 512);
  float* f0 = a.get<float>(N * 1024);
  float* f1 = a.get<float>(N * 512);
  float* f2 = a.get<float>(N * 256);
  CU(ga_concat_half(x3, g.cpos3, n, gaA16, st, &e->lc));
  CU(lin_h(e, gaA16, 264, n * 32, "ga.w1", "ga.b1", g1h, 512, 1, /*out_half=*/1, st));
  CU(lin_h(e, g1h, 512, n * 32, "ga.w2", "ga.b2", f0, 1024, 1, 0, st, nullptr, 0, /*segmax=*/1));
  CU(lin(e, true, f0, 1024, n, "lin1.w", "lin1.b", f1, 512, 1, st, nullptr, 0, 1));  // relu(lin1) (:89)
  CU(lin(e, true, f1, 512, n, "lin2.w", "lin2.b", f2, 256, 1, st, nullptr, 0, 1));   // relu(lin2) = features2 (:90)

  if (dbg) {
    if (dbg->features2) CU(cudaMemcpyAsync(dbg->features2 + static_cast<size_t>(o0) * 256, f2, N * 256 * 4, cudaMemcpyDeviceToDevice, st));
    struct { uint8_t* dst; const uint8_t* src; size_t per; } cp[] = {
        {dbg->fps1, g.fps1, 128}, {dbg->fps2, g.fps2, 64}, {dbg->fps3, g.fps3, 32}, {dbg->nbr1, g.nbr1, 128 * 32}, {dbg->nbr2, g.nbr2, 64 * 32},
        {dbg->nbr3, g.nbr3, 32 * 32}, {dbg->cnt1, g.cnt1, 128}, {dbg->cnt2, g.cnt2, 64}, {dbg->cnt3, g.cnt3, 32}};
    for (auto& c : cp)
      if (c.dst) CU(cudaMemcpyAsync(c.dst + static_cast<size_t>(o0) * c.per, c.src, N * c.per, cudaMemcpyDeviceToDevice, st));
    if (!out) return 0;
  }

  // ObjectEncoder.forward (object_encoder.py:98-149): four normalised d-dim features -> mlp_merge (d = 256 coarse, 128 fine)
  const int d = e->fine ? T2L_FINE_DIM : T2L_EMBED_DIM;
  float* cat = a.get<float>(N * 4 * d);
  float* td = a.get<float>(N * d);
  float* emb = a.get<float>(N * d);
  const float* m = meta + static_cast<size_t>(o0) * 7;
  CU(lin(e, true, f2, 256, n, "mlp_pointnet.w", "mlp_pointnet.b", td, d, 1, st));
  CU(l2_normalize_rows(td, d, cat + 0, 4L * d, n, d, st, &e->lc));
  {
    if (W(e, "color.w1").ld != 4 || W(e, "pos.w1").ld != 4 || W(e, "num.w1").ld != 4 || W(e, "color.w2").ld != 64 || W(e, "pos.w2").ld != 64 ||
        W(e, "num.w2").ld != 64 || W(e, "color.w2").rows != d)
      return fail(e, "internal: side encoder weight shapes");
    const float* w1[3] = {W(e, "color.w1").dev, W(e, "pos.w1").dev, W(e, "num.w1").dev};
    const float* b1[3] = {W(e, "color.b1").dev, W(e, "pos.b1").dev, W(e, "num.b1").dev};
    const float* w2[3] = {W(e, "color.w2").dev, W(e, "pos.w2").dev, W(e, "num.w2").dev};
    const float* b2[3] = {W(e, "color.b2").dev, W(e, "pos.b2").dev, W(e, "num.b2").dev};
    CU(side_encoders(m, n, w1, b1, w2, b2, cat, d, st, &e->lc));
  }
  CU(lin(e, true, cat, 4L * d, n, "merge.w", "merge.b", emb, d, 1, st));

  if (obj_emb_out) {
    // fine stage (models/cross_matcher.py:98-108): the per-object encodings, L2-normalised, are the result
    CU(l2_normalize_rows(emb, d, obj_emb_out + static_cast<size_t>(o0) * d, d, n, d, st, &e->lc));
    if (a.overflow) return fail(e, "internal: workspace arena too small for %d objects", n);
    return 0;
  }

  // intra-cell attention (cell_retrieval.py:85-108); fp32 throughout: these two layers amplify
  // operand rounding the most (precision table)
  float* X = a.get<float>(static_cast<size_t>(attn_rows) * 256);
  float* Xb = a.get<float>(static_cast<size_t>(attn_rows) * 256);
  float* Xbp = a.get<float>(static_cast<size_t>(attn_rows) * 512);  // hi | lo planes of the first layer's output
  float* pooled = a.get<float>(static_cast<size_t>(B) * 256);
  CU(scatter_objects_ragged(emb, cell_ptr_dev, row_ptr_dev, B, X, st, &e->lc));
  const size_t mark = a.off;
  if (cell_attention_layer(e, "obj_attn0", X, nullptr, Xb, Xbp, attn_rows, B, row_ptr_dev, cell_ptr_dev, st)) return 1;
  a.off = mark;
  if (cell_attention_layer(e, "obj_attn1", Xb, Xbp, X, nullptr, attn_rows, B, row_ptr_dev, cell_ptr_dev, st)) return 1;
  CU(max_over_rows_ragged(X, row_ptr_dev, pooled, B, st, &e->lc));
  CU(l2_normalize_rows(pooled, 256, out + static_cast<size_t>(c0) * 256, 256, B, 256, st, &e->lc));
  if (a.overflow) return fail(e, "internal: workspace arena too small for %d objects / %d cells", n, B);
  return 0;
}

static int encode_cells_impl(t2l_engine* e, const float* pts, const float* meta, const int32_t* cell_ptr, int n_cells, float* out,
                             const ObjDebug* dbg, cudaStream_t st, float* obj_emb_out = nullptr) {
  if (!e) return 1;
  if (!e->finalized) return fail(e, "weights not finalized");
  if (n_cells < 0 || !cell_ptr || cell_ptr[0] != 0) return fail(e, "encode_cells: bad cell_ptr");
  if (e->fine && !obj_emb_out && !dbg) return fail(e, "encode_cells: this engine holds the fine-stage model (CrossMatch); use t2l_fine_*");
  if (!e->fine && obj_emb_out) return fail(e, "fine_encode_objects: this engine holds the coarse model; load a CrossMatch state dict");
  ENTER_STREAM(e, st);
  for (int c = 0; c < n_cells; ++c)
    if (cell_ptr[c + 1] <= cell_ptr[c]) return fail(e, "encode_cells: cell %d has no objects (the reference asserts >= 1, cells.py:202)", c);
  int c0 = 0;
  while (c0 < n_cells) {
    int c1 = c0 + 1;
    while (c1 < n_cells && cell_ptr[c1 + 1] - cell_ptr[c0] <= e->obj_chunk) ++c1;
    if (encode_chunk(e, pts, meta, cell_ptr, c0, c1, out, dbg, st, obj_emb_out)) return 1;
    c0 = c1;
  }
  return 0;
}

extern "C" int t2l_encode_cells(t2l_engine* e, const float* pts, const float* meta, const int32_t* cell_ptr_host, int n_cells, float* out,
                                void* stream) {
  if (!pts || !meta || !out) return fail(e, "encode_cells: NULL buffer");
  if (reinterpret_cast<uintptr_t>(pts) & 15) return fail(e, "encode_cells: pts must be 16-byte aligned (bulk TMA copies read it)");
  return encode_cells_impl(e, pts, meta, cell_ptr_host, n_cells, out, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int t2l_encode_objects_debug(t2l_engine* e, const float* pts, const int32_t* cell_ptr_host, int n_cells, float* features2,
                                        uint8_t* fps1, uint8_t* fps2, uint8_t* fps3, uint8_t* nbr1, uint8_t* nbr2, uint8_t* nbr3,
                                        uint8_t* cnt1, uint8_t* cnt2, uint8_t* cnt3, void* stream) {
  ObjDebug d;
  d.features2 = features2; d.fps1 = fps1; d.fps2 = fps2; d.fps3 = fps3; d.nbr1 = nbr1; d.nbr2 = nbr2; d.nbr3 = nbr3;
  d.cnt1 = cnt1; d.cnt2 = cnt2; d.cnt3 = cnt3;
  return encode_cells_impl(e, pts, nullptr, cell_ptr_host, n_cells, nullptr, &d, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// encode_text
// ---------------------------------------------------------------------------------------------
// Token level: intra_module (one encoder layer over the tokens of each sentence, no key-padding mask,
// language_encoder.py:130-131) then max over tokens (:133).  pooled [n_seq, 1024].
static int text_tokens(t2l_engine* e, const float* t5, int n_seq, int L, float* pooled, cudaStream_t st, bool t5_is_half = false) {
  const int d = T2L_T5_DIM;
  int sc = e->tok_chunk / L;  // sentences per chunk
  if (sc < 1) sc = 1;
  if (sc > n_seq) sc = n_seq;
  if (n_seq == 0) return 0;
  if (ensure_arena(e, static_cast<size_t>(sc) * L * (12 * static_cast<size_t>(d)) * 4 + (size_t(1) << 22))) return 1;
  Arena& a = e->arena;
  for (int s0 = 0; s0 < n_seq; s0 += sc) {
    const int ns = (n_seq - s0 < sc) ? n_seq - s0 : sc;
    a.off = 0;
    const float* X = t5_is_half ? reinterpret_cast<const float*>(reinterpret_cast<const __half*>(t5) + static_cast<size_t>(s0) * L * d)
                                : t5 + static_cast<size_t>(s0) * L * d;
    if (encoder_layer(e, "txt_intra", true, X, nullptr, ns, L, d, 4 * d, st, pooled + static_cast<size_t>(s0) * d, t5_is_half)) return 1;
  }
  if (a.overflow) return fail(e, "internal: workspace arena too small for %d sentences", n_seq);
  return 0;
}

// Sentence level: inter_mlp = Linear + BN folded, no ReLU (:137), inter_module over the S sentences of
// each query with the extra residual `x += layer(x)` (:143-145), max over sentences (:147), normalise
// (cell_retrieval.py:61).
static int text_sentences(t2l_engine* e, const float* pooled, int nq, int S, float* out, cudaStream_t st) {
  const int d = T2L_T5_DIM;
  if (nq == 0) return 0;
  const size_t n_seq = static_cast<size_t>(nq) * S;
  if (ensure_arena(e, n_seq * (64 * 256) * 4 + (size_t(1) << 22))) return 1;
  Arena& a = e->arena;
  float* z = a.get<float>(n_seq * 256);
  float* z2 = a.get<float>(n_seq * 256);
  float* zq = a.get<float>(static_cast<size_t>(nq) * 256);
  CU(lin3(e, pooled, d, static_cast<int>(n_seq), "txt_mlp.w", "txt_mlp.b", z, 256, 0, st));
  if (encoder_layer(e, "txt_inter", false, z, z2, nq, S, 256, 1024, st)) return 1;
  CU(add_rows(z, z2, z2, static_cast<long>(n_seq) * 256, st, &e->lc));
  CU(max_over_rows(z2, zq, nq, S, 256, st, &e->lc));
  CU(l2_normalize_rows(zq, 256, out, 256, nq, 256, st, &e->lc));
  if (a.overflow) return fail(e, "internal: workspace arena too small for %d queries", nq);
  return 0;
}

static int text_args_ok(t2l_engine* e, const void* in, const void* out, int n, int S, int L) {
  if (!e) return 1;
  if (!e->finalized) return fail(e, "weights not finalized");
  if (e->fine) return fail(e, "encode_text: this engine holds the fine-stage model (CrossMatch); use t2l_fine_*");
  if (!in || !out || n < 0 || S < 1 || S > 32 || L < 1 || L > 32) return fail(e, "encode_text: bad argument (n_sent, n_tok must be in 1..32)");
  return 0;
}

extern "C" int t2l_encode_text_tokens(t2l_engine* e, const float* t5, int n_sentences, int L, float* pooled, void* stream) {
  if (text_args_ok(e, t5, pooled, n_sentences, 1, L)) return 1;
  ENTER_STREAM(e, stream);
  return text_tokens(e, t5, n_sentences, L, pooled, static_cast<cudaStream_t>(stream));
}

// The same with the T5 states delivered as fp16 (raw 16-bit words): half the bytes over PCIe / from HBM and no conversion
// kernel.  The token layer already runs on fp16 copies of its operands; here the out-projection's residual is read from the
// fp16 input too (one more 11-bit rounding on that operand).  Needs the fp16 token layer (the default).
extern "C" int t2l_encode_text_tokens_f16(t2l_engine* e, const void* t5_half, int n_sentences, int L, float* pooled, void* stream) {
  if (text_args_ok(e, t5_half, pooled, n_sentences, 1, L)) return 1;
  if (!e->text_f16) return fail(e, "encode_text_tokens_f16: T2L_TEXT_TF32=1 selects the tf32 token layer, which takes fp32 input");
  if (reinterpret_cast<uintptr_t>(t5_half) & 15) return fail(e, "encode_text_tokens_f16: input must be 16-byte aligned");
  ENTER_STREAM(e, stream);
  return text_tokens(e, static_cast<const float*>(t5_half), n_sentences, L, pooled, static_cast<cudaStream_t>(stream), true);
}

extern "C" int t2l_encode_text_sentences(t2l_engine* e, const float* pooled, int nq, int S, float* out, void* stream) {
  if (text_args_ok(e, pooled, out, nq, S, 1)) return 1;
  ENTER_STREAM(e, stream);
  return text_sentences(e, pooled, nq, S, out, static_cast<cudaStream_t>(stream));
}

extern "C" int t2l_encode_text(t2l_engine* e, const float* t5, int nq, int S, int L, float* out, void* stream) {
  if (text_args_ok(e, t5, out, nq, S, L)) return 1;
  ENTER_STREAM(e, stream);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n_seq = static_cast<size_t>(nq) * S;
  if (n_seq > e->pooled_cap) {
    if (e->pooled) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->pooled)); e->pooled = nullptr; }
    CU(cudaMalloc(&e->pooled, (n_seq + 1024) * T2L_T5_DIM * sizeof(float)));
    e->pooled_cap = n_seq + 1024;
  }
  if (text_tokens(e, t5, static_cast<int>(n_seq), L, e->pooled, st)) return 1;
  return text_sentences(e, e->pooled, nq, S, out, st);
}


// ---------------------------------------------------------------------------------------------
// fine stage: CrossMatch.forward (models/cross_matcher.py:83-129), SURVEY.md section 8f row 1 / BASELINE configs[4]
// ---------------------------------------------------------------------------------------------
// One post-norm nn.TransformerDecoderLayer (ReLU, eps 1e-5, eval, no masks; cross_matcher.py:66-72) on packed rows:
//   x = LN1(t + SelfAttn(t));  x = LN2(x + CrossAttn(x, mem));  x = LN3(x + FFN(x))
// T [n_seq * St, d], Mem [n_seq * Sm, d] -> Tout [n_seq * St, d].  All projections are three-pass split products (fp32 accuracy).
static int decoder_layer(t2l_engine* e, const std::string& pfx, const float* T, const float* Tp, int St, const float* Mem, const float* Memp, int Sm,
                         float* Tout, float* Toutp, int n_seq, int d, cudaStream_t st) {
  // Tp / Memp / Toutp: tf32 hi | lo planes ([rows, 2d]) of the target, the memory and the output.  Every three-pass GEMM of the
  // layer reads planes its producer wrote (LayerNorms, attention cores, the ReLU epilogue): no split kernels inside the layer.
  const int rt = n_seq * St, rm = n_seq * Sm;
  Arena& a = e->arena;
  const size_t mark = a.off;
  float* qkv = a.get<float>(static_cast<size_t>(rt) * 3 * d);
  float* attp = a.get<float>(static_cast<size_t>(rt) * 2 * d);
  float* y = a.get<float>(static_cast<size_t>(rt) * d);
  float* x1 = a.get<float>(static_cast<size_t>(rt) * d);
  float* x1p = a.get<float>(static_cast<size_t>(rt) * 2 * d);
  float* q2 = a.get<float>(static_cast<size_t>(rt) * d);
  float* kv2 = a.get<float>(static_cast<size_t>(rm) * 2 * d);
  float* x2 = a.get<float>(static_cast<size_t>(rt) * d);
  float* x2p = a.get<float>(static_cast<size_t>(rt) * 2 * d);
  float* hp = a.get<float>(static_cast<size_t>(rt) * 8 * d);
  CU(lin3(e, T, d, rt, pfx + ".in_w", pfx + ".in_b", qkv, 3L * d, 0, st, nullptr, 0, Tp));
  CU(mha_small(qkv, nullptr, n_seq, St, d, 4, st, &e->lc, 0, attp));
  CU(lin3(e, nullptr, d, rt, pfx + ".out_w", pfx + ".out_b", y, d, 0, st, T, d, attp));
  CU(layer_norm_rows(y, x1, W(e, pfx + ".n1_w").dev, W(e, pfx + ".n1_b").dev, rt, d, st, &e->lc, nullptr, x1p));
  CU(lin3(e, nullptr, d, rt, pfx + ".ca_q_w", pfx + ".ca_q_b", q2, d, 0, st, nullptr, 0, x1p));
  CU(lin3(e, Mem, d, rm, pfx + ".ca_kv_w", pfx + ".ca_kv_b", kv2, 2L * d, 0, st, nullptr, 0, Memp));
  CU(mha_cross_small(q2, d, kv2, kv2 + d, 2L * d, nullptr, n_seq, St, Sm, d, 4, st, &e->lc, 0, attp));
  CU(lin3(e, nullptr, d, rt, pfx + ".ca_out_w", pfx + ".ca_out_b", y, d, 0, st, x1, d, attp));
  CU(layer_norm_rows(y, x2, W(e, pfx + ".n2_w").dev, W(e, pfx + ".n2_b").dev, rt, d, st, &e->lc, nullptr, x2p));
  CU(lin3(e, nullptr, d, rt, pfx + ".l1_w", pfx + ".l1_b", hp, 8L * d, 1, st, nullptr, 0, x2p, /*split_out=*/1));
  CU(lin3(e, nullptr, 4 * d, rt, pfx + ".l2_w", pfx + ".l2_b", y, d, 0, st, x2, d, hp));
  CU(layer_norm_rows(y, Tout, W(e, pfx + ".n3_w").dev, W(e, pfx + ".n3_b").dev, rt, d, st, &e->lc, nullptr, Toutp));
  a.off = mark;  // Tout / Toutp live outside the scratch of this layer
  return 0;
}

static int fine_args_ok(t2l_engine* e) {
  if (!e) return 1;
  if (!e->finalized) return fail(e, "weights not finalized");
  if (!e->fine) return fail(e, "this engine holds the coarse model; the fine stage needs a CrossMatch state dict");
  return 0;
}

extern "C" int t2l_fine_encode_objects(t2l_engine* e, const float* pts, const float* meta, const int32_t* cell_ptr_host, int n_cells,
                                       float* obj_emb, void* stream) {
  if (fine_args_ok(e)) return 1;
  if (!pts || !meta || !obj_emb) return fail(e, "fine_encode_objects: NULL buffer");
  if (reinterpret_cast<uintptr_t>(pts) & 15) return fail(e, "fine_encode_objects: pts must be 16-byte aligned");
  return encode_cells_impl(e, pts, meta, cell_ptr_host, n_cells, nullptr, nullptr, static_cast<cudaStream_t>(stream), obj_emb);
}

// LanguageEncoder(is_fine=True) after T5 (models/language_encoder.py:130-140): token layer, max over tokens, inter_mlp
static int fine_hints_impl(t2l_engine* e, const float* t5, int n_sent, int L, float* hints, cudaStream_t st) {
  const int d = T2L_FINE_DIM;
  if (n_sent == 0) return 0;
  if (static_cast<size_t>(n_sent) > e->pooled_cap) {
    if (e->pooled) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->pooled)); e->pooled = nullptr; }
    CU(cudaMalloc(&e->pooled, (static_cast<size_t>(n_sent) + 1024) * T2L_T5_DIM * sizeof(float)));
    e->pooled_cap = static_cast<size_t>(n_sent) + 1024;
  }
  if (text_tokens(e, t5, n_sent, L, e->pooled, st)) return 1;
  if (ensure_arena(e, static_cast<size_t>(n_sent) * 2 * T2L_T5_DIM * 4 + (size_t(1) << 22))) return 1;
  CU(lin3(e, e->pooled, T2L_T5_DIM, n_sent, "txt_mlp.w", "txt_mlp.b", hints, d, 0, st));
  if (e->arena.overflow) return fail(e, "internal: workspace arena too small for %d hint sentences", n_sent);
  return 0;
}

extern "C" int t2l_fine_encode_hints(t2l_engine* e, const float* t5, int n_sentences, int n_tok, float* hints, void* stream) {
  if (fine_args_ok(e)) return 1;
  if (!t5 || !hints || n_sentences < 0 || n_tok < 1 || n_tok > 32) return fail(e, "fine_encode_hints: bad argument (n_tok in 1..32)");
  ENTER_STREAM(e, stream);
  return fine_hints_impl(e, t5, n_sentences, n_tok, hints, static_cast<cudaStream_t>(stream));
}

static int fine_match_impl(t2l_engine* e, const float* obj_emb, const int32_t* pair_cell, const float* hints, const int32_t* pair_query, int n_pairs,
                           int n_obj, int n_hints, float* offsets, cudaStream_t st) {
  const int d = T2L_FINE_DIM;
  const int chunk = 16384;  // pairs per pass (~300 KB of scratch each; 4 096 -> 16 384: a quarter of the ~100 small launches per pass)
  for (int p0 = 0; p0 < n_pairs; p0 += chunk) {
    const int np = n_pairs - p0 < chunk ? n_pairs - p0 : chunk;
    if (ensure_arena(e, static_cast<size_t>(np) * (static_cast<size_t>(n_obj) + n_hints) * d * 4 * 40 + (size_t(1) << 22))) return 1;
    Arena& a = e->arena;
    float* d0 = a.get<float>(static_cast<size_t>(np) * n_obj * d);
    float* d0b = a.get<float>(static_cast<size_t>(np) * n_obj * d);
    float* d1 = a.get<float>(static_cast<size_t>(np) * n_hints * d);
    float* d1b = a.get<float>(static_cast<size_t>(np) * n_hints * d);
    // tf32 hi | lo planes of the four row sets (operands of the decoder layers' three-pass GEMMs)
    float* d0p = a.get<float>(static_cast<size_t>(np) * n_obj * 2 * d);
    float* d0bp = a.get<float>(static_cast<size_t>(np) * n_obj * 2 * d);
    float* d1p = a.get<float>(static_cast<size_t>(np) * n_hints * 2 * d);
    float* d1bp = a.get<float>(static_cast<size_t>(np) * n_hints * 2 * d);
    float* hmax = a.get<float>(static_cast<size_t>(np) * d);
    float* h64 = a.get<float>(static_cast<size_t>(np) * (d / 2));
    // desc0 / desc1 of every pair: rows of the cell's (already normalised) objects, rows of the query's hints (:105-111)
    CU(gather_row_groups(obj_emb, pair_cell ? pair_cell + p0 : nullptr, p0, np, n_obj, d, d0, st, &e->lc));
    CU(gather_row_groups(hints, pair_query ? pair_query + p0 : nullptr, p0, np, n_hints, d, d1, st, &e->lc));
    // cascaded cross-attention (:113-115): objects attend to hints, then hints to the updated objects, twice
    CU(split_tf32_planes(d0, d, d0p, np * n_obj, d, st, &e->lc));
    CU(split_tf32_planes(d1, d, d1p, np * n_hints, d, st, &e->lc));
    float *o_in = d0, *o_out = d0b, *h_in = d1, *h_out = d1b;
    float *o_inp = d0p, *o_outp = d0bp, *h_inp = d1p, *h_outp = d1bp;
    for (int i = 0; i < 2; ++i) {
      if (decoder_layer(e, "cross_objects" + std::to_string(i), o_in, o_inp, n_obj, h_in, h_inp, n_hints, o_out, o_outp, np, d, st)) return 1;
      if (decoder_layer(e, "cross_hints" + std::to_string(i), h_in, h_inp, n_hints, o_out, o_outp, n_obj, h_out, h_outp, np, d, st)) return 1;
      std::swap(o_in, o_out);
      std::swap(h_in, h_out);
      std::swap(o_inp, o_outp);
      std::swap(h_inp, h_outp);
    }
    CU(max_over_rows(h_in, hmax, np, n_hints, d, st, &e->lc));  // desc1.max(dim=0) (:126)
    // mlp_offsets = Linear(d, d/2) + ReLU + Linear(d/2, 2) (:17-36, :127), exact fp32
    CU(lin(e, false, hmax, d, np, "offs.w1", "offs.b1", h64, d / 2, 1, st));
    CU(lin(e, false, h64, d / 2, np, "offs.w2", "offs.b2", offsets + static_cast<size_t>(p0) * 2, 2, 0, st));
    if (a.overflow) return fail(e, "internal: workspace arena too small for %d pairs", np);
  }
  return 0;
}

extern "C" int t2l_fine_match(t2l_engine* e, const float* obj_emb, const int32_t* pair_cell, const float* hints, const int32_t* pair_query,
                              int n_pairs, int n_obj, int n_hints, float* offsets, void* stream) {
  if (fine_args_ok(e)) return 1;
  if (!obj_emb || !hints || !offsets || n_pairs < 0 || n_obj < 1 || n_obj > 32 || n_hints < 1 || n_hints > 32)
    return fail(e, "fine_match: bad argument (objects per cell and hints per query in 1..32)");
  ENTER_STREAM(e, stream);
  return fine_match_impl(e, obj_emb, pair_cell, hints, pair_query, n_pairs, n_obj, n_hints, offsets, static_cast<cudaStream_t>(stream));
}

extern "C" int t2l_fine_offsets(t2l_engine* e, const float* pts, const float* meta, const int32_t* cell_ptr_host, int n_cells, const float* t5,
                                int n_hints, int n_tok, float* offsets, void* stream) {
  if (fine_args_ok(e)) return 1;
  if (!pts || !meta || !cell_ptr_host || !t5 || !offsets || n_cells < 0 || n_hints < 1 || n_hints > 32 || n_tok < 1 || n_tok > 32)
    return fail(e, "fine_offsets: bad argument");
  if (n_cells == 0) return 0;
  const int n_obj = cell_ptr_host[1] - cell_ptr_host[0];
  for (int c = 0; c < n_cells; ++c)
    if (cell_ptr_host[c + 1] - cell_ptr_host[c] != n_obj)
      return fail(e, "fine_offsets: every cell must hold the same number of (padded) objects (pad_size, dataloading/kitti360pose/eval.py:147-160)");
  if (n_obj < 1 || n_obj > 32) return fail(e, "fine_offsets: objects per cell must be in 1..32");
  const int d = T2L_FINE_DIM;
  const size_t need = static_cast<size_t>(n_cells) * (static_cast<size_t>(n_obj) + n_hints) * d;
  if (need > e->fine_cap) {
    ENTER(e);
    if (e->fine_buf) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->fine_buf)); e->fine_buf = nullptr; }
    CU(cudaMalloc(&e->fine_buf, (need + 4096) * sizeof(float)));
    e->fine_cap = need + 4096;
  }
  float* obj_emb = e->fine_buf;
  float* hints = e->fine_buf + static_cast<size_t>(n_cells) * n_obj * d;
  if (t2l_fine_encode_objects(e, pts, meta, cell_ptr_host, n_cells, obj_emb, stream)) return 1;
  if (t2l_fine_encode_hints(e, t5, n_cells * n_hints, n_tok, hints, stream)) return 1;
  return t2l_fine_match(e, obj_emb, nullptr, hints, nullptr, n_cells, n_obj, n_hints, offsets, stream);
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
static int ensure_search_work(t2l_engine* e, int nq);

// Size the workspace once for the largest calls to come, so that no later call re-allocates (and synchronises).
extern "C" int t2l_reserve(t2l_engine* e, int max_objects, int max_cells, int max_sentences, int max_tokens_per_sentence, int max_queries) {
  if (!e) return 1;
  if (max_objects < 0 || max_cells < 0 || max_sentences < 0 || max_tokens_per_sentence < 0 || max_queries < 0) return fail(e, "reserve: bad argument");
  ENTER(e);
  size_t need = 0;
  if (max_objects > 0) {
    const size_t n = static_cast<size_t>(max_objects < e->obj_chunk ? max_objects : e->obj_chunk);
    need = obj_chunk_bytes(n, static_cast<size_t>(max_cells < static_cast<int>(n) ? max_cells : static_cast<int>(n)));
  }
  if (max_sentences > 0 && max_tokens_per_sentence > 0) {
    int sc = e->tok_chunk / max_tokens_per_sentence;
    if (sc < 1) sc = 1;
    if (sc > max_sentences) sc = max_sentences;
    const size_t t = static_cast<size_t>(sc) * max_tokens_per_sentence * (12 * static_cast<size_t>(T2L_T5_DIM)) * 4 + (size_t(1) << 22);
    if (t > need) need = t;
    const size_t s2 = static_cast<size_t>(max_sentences) * (64 * 256) * 4 + (size_t(1) << 22);
    if (s2 > need) need = s2;
    if (static_cast<size_t>(max_sentences) > e->pooled_cap) {
      if (e->pooled) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->pooled)); e->pooled = nullptr; }
      CU(cudaMalloc(&e->pooled, (static_cast<size_t>(max_sentences) + 1024) * T2L_T5_DIM * sizeof(float)));
      e->pooled_cap = static_cast<size_t>(max_sentences) + 1024;
    }
  }
  if (need && ensure_arena(e, need)) return 1;
  if (max_queries > 0 && ensure_search_work(e, max_queries)) return 1;
  return 0;
}

extern "C" int t2l_db_build(t2l_engine* e, const float* D, int64_t n_rows, int64_t row_offset, void* stream) {
  if (!e) return 1;
  if (n_rows < 0 || (n_rows > 0 && !D) || n_rows > 0x7fffff00LL) return fail(e, "db_build: bad argument");
  ENTER_STREAM(e, stream);
  if (static_cast<size_t>(n_rows) > e->sw_planes_rows) {
    if (e->db.planes) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->db.planes)); CU(cudaFree(e->db.plane16)); e->db.planes = nullptr; e->db.plane16 = nullptr; }
    CU(cudaMalloc(&e->db.planes, static_cast<size_t>(n_rows) * 512 * sizeof(__nv_bfloat16) + 1024));
    CU(cudaMalloc(&e->db.plane16, static_cast<size_t>(n_rows) * 256 * sizeof(__half) + 1024));
    e->sw_planes_rows = static_cast<size_t>(n_rows);
  }
  e->db.D = D; e->db.n_rows = n_rows; e->db.row_offset = row_offset;
  CU(search_prepare_db(e->db, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

static int ensure_search_work(t2l_engine* e, int nq) {
  if (nq <= e->sw.nq_cap) return 0;
  CU(cudaDeviceSynchronize());
  free_search_work(e);
  const size_t cap = static_cast<size_t>(nq) + 128;
  const int sc = 16;
  CU(cudaMalloc(&e->sw.q_planes, cap * 512 * sizeof(__nv_bfloat16)));
  CU(cudaMalloc(&e->sw.q16, cap * 256 * sizeof(__half)));
  CU(cudaMalloc(&e->sw.q_scale, cap * sizeof(float)));
  CU(cudaMalloc(&e->sw.q_norm, cap * sizeof(float)));
  CU(cudaMalloc(&e->sw.cand_score, cap * sc * 16 * sizeof(float)));
  CU(cudaMalloc(&e->sw.cand_idx, cap * sc * 16 * sizeof(int32_t)));
  CU(cudaMalloc(&e->sw.cand_thr, cap * sc * sizeof(float)));
  CU(cudaMalloc(&e->sw.flags, cap * sizeof(int32_t)));
  CU(cudaMalloc(&e->sw.n_fail, sizeof(int32_t)));
  CU(cudaMalloc(&e->sw.fail_ids, cap * sizeof(int32_t)));
  CU(cudaMalloc(&e->sw.fail_thr, cap * sizeof(float)));
  CU(cudaMalloc(&e->sw.q2_planes, cap * 512 * sizeof(__nv_bfloat16)));
  CU(cudaMalloc(&e->sw.cand2_idx, cap * kPass2Cap * sizeof(int32_t)));
  CU(cudaMalloc(&e->sw.cand2_cnt, cap * sc * sizeof(int32_t)));
  e->sw.nq_cap = static_cast<int>(cap);
  e->sw.splits_cap = sc;
  return 0;
}

extern "C" int t2l_search_topk(t2l_engine* e, const float* Q, int nq, int k, int64_t* out_idx, double* out_score, int32_t* out_n_fallback,
                               void* stream) {
  if (!e) return 1;
  if (!Q || !out_idx || !out_score || nq < 0) return fail(e, "search_topk: bad argument");
  if (k < 1 || k > T2L_MAX_TOPK) return fail(e, "search_topk: k must be in 1..%d (use t2l_search_topk_exact beyond)", T2L_MAX_TOPK);
  if (!e->db.D && e->db.n_rows != 0) return fail(e, "search_topk: t2l_db_build has not been called");
  ENTER_STREAM(e, stream);
  if (ensure_search_work(e, nq)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CU(cudaStreamIsCapturing(st, &cap));
  const bool capturing = cap != cudaStreamCaptureStatusNone;  // (a captured call keeps the mode it was captured with)
  if (e->search_first == 0 && !capturing) {
    if (e->fail_pending) {
      const cudaError_t q = cudaEventQuery(e->fail_ev);
      if (q == cudaSuccess) {
        e->fail_pending = false;
        if (e->fail_nq >= 64 && *e->fail_host > 0.30 * e->fail_nq) { e->bf16_first_now = true; e->bf16_first_calls = 0; }
      } else {
        (void)cudaGetLastError();  // cudaErrorNotReady is not an error; keep it out of the launch checks
      }
    }
    if (e->bf16_first_now && ++e->bf16_first_calls > 16) e->bf16_first_now = false;  // probe the fp16 pass again now and then
  }
  const bool bf16_first = e->search_first == 2 || (e->search_first == 0 && e->bf16_first_now);
  CU(search_topk(e->db, e->sw, Q, nq, k, out_idx, out_score, out_n_fallback, bf16_first, st, &e->lc));
  if (e->search_first == 0 && !capturing && !bf16_first && !e->fail_pending && nq > 0 && e->db.n_rows > 0) {
    CU(cudaMemcpyAsync(e->fail_host, e->sw.n_fail, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(e->fail_ev, st));
    e->fail_pending = true;
    e->fail_nq = nq;
  }
  return 0;
}

// Search the registered shard and fold its top-k into the caller's running lists (streamed databases: encode a chunk,
// db_build it with its row offset, accumulate).
extern "C" int t2l_search_topk_accumulate(t2l_engine* e, const float* Q, int nq, int k, int64_t* run_idx, double* run_score,
                                          int32_t* out_n_fallback, void* stream) {
  if (!e) return 1;
  if (!run_idx || !run_score) return fail(e, "search_topk_accumulate: NULL running list");
  if (nq > e->acc_cap) {
    ENTER(e);
    if (e->acc_idx) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->acc_idx)); CU(cudaFree(e->acc_score)); e->acc_idx = nullptr; e->acc_score = nullptr; }
    CU(cudaMalloc(&e->acc_idx, (static_cast<size_t>(nq) + 128) * T2L_MAX_TOPK * sizeof(int64_t)));
    CU(cudaMalloc(&e->acc_score, (static_cast<size_t>(nq) + 128) * T2L_MAX_TOPK * sizeof(double)));
    e->acc_cap = nq + 128;
  }
  if (t2l_search_topk(e, Q, nq, k, e->acc_idx, e->acc_score, out_n_fallback, stream)) return 1;
  ENTER_STREAM(e, stream);
  CU(merge_running_topk(run_idx, run_score, e->acc_idx, e->acc_score, nq, k, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

extern "C" int t2l_synth_cells(t2l_engine* e, uint64_t seed, int64_t first_cell, int n_cells, int obj_per_cell, float* pts, float* meta,
                               void* stream) {
  if (!e) return 1;
  if (!pts || !meta || n_cells < 0 || obj_per_cell < 1 || first_cell < 0) return fail(e, "synth_cells: bad argument");
  ENTER_STREAM(e, stream);
  CU(synth_cells(seed, first_cell * obj_per_cell, static_cast<long>(n_cells) * obj_per_cell, pts, meta, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

extern "C" int t2l_search_topk_exact(t2l_engine* e, const float* Q, int nq, int k, int64_t* out_idx, double* out_score, void* stream) {
  if (!e) return 1;
  if (!Q || !out_idx || !out_score || nq < 0 || k < 1 || k > 16) return fail(e, "search_topk_exact: bad argument (k in 1..16)");
  ENTER_STREAM(e, stream);
  CU(search_topk_exact(e->db, Q, nq, k, out_idx, out_score, nullptr, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

extern "C" int t2l_merge_topk(t2l_engine* e, const int64_t* idx_all, const double* score_all, int n_shards, int nq, int k, int64_t* out_idx,
                              double* out_score, void* stream) {
  if (!e) return 1;
  if (!idx_all || !score_all || !out_idx || !out_score) return fail(e, "merge_topk: NULL buffer");
  ENTER_STREAM(e, stream);
  CU(merge_topk(idx_all, score_all, static_cast<long>(nq) * k, n_shards, nq, k, out_idx, out_score, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

// The same for PACKED per-shard results: shard g contributes one block of 2 * nq * k 8-byte words, [idx i64 [nq, k] | score f64
// [nq, k]] -- what ONE all-gather of a rank's (idx, score) pair delivers (text2loc_b200/distributed.py).
extern "C" int t2l_merge_topk_packed(t2l_engine* e, const void* packed_all, int n_shards, int nq, int k, int64_t* out_idx, double* out_score,
                                     void* stream) {
  if (!e) return 1;
  if (!packed_all || !out_idx || !out_score) return fail(e, "merge_topk_packed: NULL buffer");
  ENTER_STREAM(e, stream);
  const int64_t* idx0 = static_cast<const int64_t*>(packed_all);
  const double* sc0 = reinterpret_cast<const double*>(idx0 + static_cast<long>(nq) * k);
  CU(merge_topk(idx0, sc0, 2L * nq * k, n_shards, nq, k, out_idx, out_score, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

extern "C" int t2l_topk_accuracy(t2l_engine* e, const int64_t* idx, int nq, int k, const int64_t* target_row, const double* query_xy,
                                 const double* cell_xy, const int32_t* query_scene, const int32_t* cell_scene, const int32_t* top_k_host,
                                 int n_top, const double* threshs_host, int n_thr, uint8_t* hit, uint8_t* within, double* dists,
                                 void* stream) {
  if (!e) return 1;
  if (!idx || !query_xy || !cell_xy || nq < 0 || k < 1 || n_top < 0 || n_top > 8 || n_thr < 0 || n_thr > 8 || (n_top && !top_k_host) ||
      (n_thr && !threshs_host) || (within && !n_thr) || ((cell_scene != nullptr) != (query_scene != nullptr)))
    return fail(e, "topk_accuracy: bad argument (at most 8 k values and 8 thresholds)");
  TopkAccuracy a{};
  a.idx = idx; a.nq = nq; a.k = k; a.target_row = target_row; a.query_xy = query_xy; a.cell_xy = cell_xy;
  a.query_scene = query_scene; a.cell_scene = cell_scene; a.n_top = n_top; a.n_thr = n_thr; a.hit = hit; a.within = within; a.dists = dists;
  for (int i = 0; i < n_top; ++i) {
    if (top_k_host[i] < 1 || (i && top_k_host[i] <= top_k_host[i - 1])) return fail(e, "topk_accuracy: top_k must be ascending and >= 1");
    a.top_k[i] = top_k_host[i];
  }
  for (int i = 0; i < n_thr; ++i) a.threshs[i] = threshs_host[i];
  ENTER_STREAM(e, stream);
  CU(topk_accuracy(a, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// test hook
// ---------------------------------------------------------------------------------------------
extern "C" int t2l_debug_linear(t2l_engine* e, int path, const float* A, int lda, const float* Wt, int ldw, const float* bias, float* C, int ldc,
                                int M, int N, int K, int act, int segmax, void* stream) {
  if (!e) return 1;
  ENTER_STREAM(e, stream);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Linear l;
  l.A = A; l.lda = lda; l.W = Wt; l.ldw = ldw; l.bias = bias; l.C = C; l.ldc = ldc; l.M = M; l.N = N; l.K = K; l.act = act;
  if (path == 1) {
    l.segmax = segmax;
    CU(linear_umma(l, st, &e->lc));
    return 0;
  }
  if (!segmax) { CU(linear_simt(l, st, &e->lc)); return 0; }
  if (ensure_arena(e, static_cast<size_t>(M) * N * 4 + 4096)) return 1;
  float* tmp = e->arena.get<float>(static_cast<size_t>(M) * N);
  l.C = tmp; l.ldc = N; l.act = 1;
  CU(linear_simt(l, st, &e->lc));
  CU(segmax32(tmp, N, C, ldc, nullptr, 0, M / 32, N, st, &e->lc));
  return 0;
}

extern "C" int t2l_debug_linear_f16(t2l_engine* e, const void* A, int lda, const void* Wt, int ldw, const float* bias, void* C, int ldc,
                                    int M, int N, int K, int act, int out_half, void* stream) {
  if (!e) return 1;
  ENTER_STREAM(e, stream);
  Linear l;
  l.A = static_cast<const float*>(A); l.lda = lda; l.W = static_cast<const float*>(Wt); l.ldw = ldw; l.bias = bias;
  l.C = static_cast<float*>(C); l.ldc = ldc; l.M = M; l.N = N; l.K = K; l.act = act; l.half_ops = 1; l.out_half = out_half;
  CU(linear_umma(l, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

extern "C" int t2l_debug_linear_f16_residual(t2l_engine* e, const void* A, int lda, const void* Wt, int ldw, const float* bias, const void* R,
                                             int ldr, void* C, int ldc, int M, int N, int K, int reg_epilogue, void* stream) {
  if (!e) return 1;
  if (!A || !Wt || !R || !C) return fail(e, "debug_linear_f16_residual: NULL buffer");
  ENTER_STREAM(e, stream);
  Linear l;
  l.A = static_cast<const float*>(A); l.lda = lda; l.W = static_cast<const float*>(Wt); l.ldw = ldw; l.bias = bias;
  l.C = static_cast<float*>(C); l.ldc = ldc; l.M = M; l.N = N; l.K = K; l.half_ops = 1; l.out_half = 1;
  l.residual = static_cast<const float*>(R); l.ldr = ldr; l.residual_half = 1; l.reg_epilogue = reg_epilogue;
  CU(linear_umma(l, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

extern "C" int t2l_debug_mha(t2l_engine* e, const float* qkv, float* out, int n_seq, int S, int d, int n_heads, void* stream) {
  if (!e) return 1;
  if (!qkv || !out || n_seq < 0 || S < 1 || S > 32 || n_heads < 1 || d % n_heads) return fail(e, "debug_mha: bad argument");
  ENTER_STREAM(e, stream);
  CU(mha_small(qkv, out, n_seq, S, d, n_heads, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

extern "C" int t2l_debug_mha_cells(t2l_engine* e, const float* qkv, float* out, int n_cells, const int32_t* row_ptr_dev,
                                   const int32_t* cell_ptr_dev, int slots, int d, int n_heads, void* stream) {
  if (!e) return 1;
  if (!qkv || !out || !row_ptr_dev || !cell_ptr_dev || n_cells < 0) return fail(e, "debug_mha_cells: bad argument");
  ENTER_STREAM(e, stream);
  CU(mha_cells64(qkv, out, n_cells, row_ptr_dev, cell_ptr_dev, slots, d, n_heads, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}

extern "C" int t2l_debug_sa_bisect(t2l_engine* e, int mode) {
  if (!e) return 1;
  if (mode < 0 || mode > 3) return fail(e, "debug_sa_bisect: mode must be 0..3");
  e->sa_bisect = mode;
  return 0;
}

extern "C" int t2l_debug_mha_cross(t2l_engine* e, const float* q, const float* kv, float* out, int n_seq, int Sq, int Sk, int d, int n_heads,
                                   void* stream) {
  if (!e) return 1;
  if (!q || !kv || !out || n_seq < 0 || Sq < 1 || Sk < 1 || Sk > 32 || n_heads < 1 || d % n_heads) return fail(e, "debug_mha_cross: bad argument");
  ENTER_STREAM(e, stream);
  CU(mha_cross_small(q, d, kv, kv + d, 2L * d, out, n_seq, Sq, Sk, d, n_heads, static_cast<cudaStream_t>(stream), &e->lc));
  return 0;
}
