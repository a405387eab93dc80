// mcgaze_b200 engine + C ABI (include/mcgaze_b200.h).
//
// Host-side runtime for the MCGaze per-clip forward: checkpoint ingestion (BN folding, K-major
// repack, split-fp16), workspace arena, launch schedule of the trunk (ResNet-50 + FPN) and of the
// 4-stage query head, intermediates registry for per-op parity tests, optional CUDA-graph replay.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mcgaze_b200.h"
#include "common.cuh"
#include "head_kernels.cuh"
#include "simt_gemm.cuh"
#include "umma_gemm.cuh"
#include "bneck_fused.cuh"
#include "stem_fused.cuh"

namespace mcg {

static thread_local std::string g_last_error;

// ------------------------------------------------------------------------------------ memory
struct DeviceBlock {
  void* p = nullptr;
  size_t bytes = 0;
  DeviceBlock() = default;
  explicit DeviceBlock(size_t n) : bytes(n) { MCG_CUDA(cudaMalloc(&p, n ? n : 16)); }
  DeviceBlock(const DeviceBlock&) = delete;
  DeviceBlock& operator=(const DeviceBlock&) = delete;
  ~DeviceBlock() {
    if (p) cudaFree(p);
  }
};

// Bump allocator over one cudaMalloc; reset when the forward shape changes.
class Arena {
 public:
  // measuring pass: alloc() only adds up sizes (returns a fake, never dereferenced pointer)
  void begin_measure() {
    measuring_ = true;
    off_ = 0;
  }
  size_t end_measure() {
    measuring_ = false;
    const size_t n = off_;
    off_ = 0;
    return n;
  }
  // grow-only; reserve(0) just rewinds
  void reserve(size_t bytes) {
    if (bytes > cap_) {
      block_.reset();
      block_.reset(new DeviceBlock(bytes));
      cap_ = bytes;
    }
    off_ = 0;
  }
  size_t capacity() const { return cap_; }
  template <typename T>
  T* alloc(size_t count) {
    const size_t bytes = (count * sizeof(T) + 1023) & ~static_cast<size_t>(1023);
    if (measuring_) {
      off_ += bytes;
      return reinterpret_cast<T*>(static_cast<uintptr_t>(1024));
    }
    MCG_CHECK(off_ + bytes <= cap_, "arena overflow");
    T* r = reinterpret_cast<T*>(static_cast<uint8_t*>(block_->p) + off_);
    off_ += bytes;
    return r;
  }
  size_t used() const { return off_; }

 private:
  std::unique_ptr<DeviceBlock> block_;
  size_t cap_ = 0, off_ = 0;
  bool measuring_ = false;
};

template <typename T>
static T* upload(std::vector<std::unique_ptr<DeviceBlock>>& keep, const std::vector<T>& host) {
  keep.emplace_back(new DeviceBlock(host.size() * sizeof(T)));
  MCG_CUDA(cudaMemcpy(keep.back()->p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return reinterpret_cast<T*>(keep.back()->p);
}

static inline __half h_from_float(float v) { return __float2half_rn(v); }

// e4m3 weight planes of the fp16c8 mode: hi8 = e4m3(hi 2^kW8HiShift), lo8 = e4m3((w - hi) 2^kW8LoShift)
static void pack_w8_host(const float* w, const __half* hi, size_t n, uint8_t* hi8, uint8_t* lo8) {
  const float s1 = std::ldexp(1.f, kW8HiShift), s2 = std::ldexp(1.f, kW8LoShift);
  for (size_t i = 0; i < n; ++i) {
    const float h = __half2float(hi[i]);
    hi8[i] = float_to_e4m3(h * s1);
    lo8[i] = float_to_e4m3((w[i] - h) * s2);
  }
}

// ------------------------------------------------------------------------------------ weights
struct GemmW {  // packed [N, K] weight (+ bias) on device
  int N = 0, K = 0;
  const float* w_f32 = nullptr;
  const float* w_t = nullptr;  // [K, N] transposed copy for small_linear_kernel (small layers only)
  Planes w;                    // hi / lo fp16 planes; fp16c8 conv weights also carry w.hi8 = e4m3(hi 2^4),
                               // w.lo8 = e4m3(lo 2^15)
  const float* bias = nullptr;
};
struct LnW {
  const float* g = nullptr;
  const float* b = nullptr;
  int C = 0;
};
struct ConvW {
  GemmW g;
  int Cin = 0, Cout = 0, R = 1, S = 1, stride = 1, pad = 0;
};
struct BlockW {
  ConvW c1, c2, c3, ds;
  ConvW c3ds;  // conv3 and the downsample branch K-concatenated: [Cout, planes + Cin], bias = b3 + b_ds
  bool has_ds = false;
};
struct StageW {
  GemmW in_proj, out_proj, dyn, fc, ffn1, ffn2, cls_fc, reg_fc[3], fc_cls[3], fc_reg[3];
  LnW attn_norm, norm_in, norm_out, fc_norm, iic_norm, ffn_norm, cls_ln, reg_ln[3];
};
struct GazeW {
  GemmW tower[3][2], ctower[3][2], fc[3], fc_conf[3];
  LnW tower_ln[3][2], ctower_ln[3][2];
  const float* wg = nullptr;
  const float* bg = nullptr;
};

constexpr int kFcSplit = 14;  // 196 k-blocks -> 14 per slice
constexpr int kFfnSplit = 4;  // second FFN Linear (K = 2048): 32 k-blocks -> 8 per slice, 24 -> 96 tiles
constexpr int kDefaultSplitLayers = 0;  // Engine::split_layers_: bit l = layer l+1 runs as two half-batch chains

struct Interm {
  int kind = 0;  // 0 = fp32 dense, 1 = NHWC planes
  const float* f32 = nullptr;
  Planes pl;
  int64_t shape[4] = {0, 0, 0, 0};  // kind 1: NB,H,W,C
};

// one row of mcg_range_report: value-range statistics of a tensor against the e4m3 windows of the fp16c8 mode
struct RangeRow {
  std::string name;
  unsigned long long total = 0, nonzero = 0, over = 0, under = 0, nonfinite = 0;
  float maxabs = 0.f;
  double energy = 0.0, energy_under = 0.0;
};

// ------------------------------------------------------------------------------------ engine
class Engine {
 public:
  Engine(int device, const mcg_tensor* w, int n, int precision) : device_(device), precision_(precision) {
    MCG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MCG_CUDA(cudaGetDeviceProperties(&prop, device));
    num_sms_ = prop.multiProcessorCount;
    if (precision != MCG_PRECISION_SIMT)
      MCG_CHECK(prop.major == 10, "tcgen05 path needs an sm_100 device, found sm_" + std::to_string(prop.major) +
                                      std::to_string(prop.minor));
    for (int i = 0; i < n; ++i) {
      size_t cnt = 1;
      for (int d = 0; d < w[i].ndim; ++d) cnt *= static_cast<size_t>(w[i].shape[d]);
      host_[w[i].name] = {w[i].data, cnt};
    }
    // DynamicConv on the warp-level tensor cores (needs the permuted dynamic_layer rows); the fp32 CUDA-core mode and
    // MCG_TUNE_DYNCONV_FFMA=1 keep the FFMA kernel and the reference's row order
    dyn_mma_ = precision != MCG_PRECISION_SIMT && std::getenv("MCG_TUNE_DYNCONV_FFMA") == nullptr;
    load_weights();
    host_.clear();
    MCG_CUDA(cudaFuncSetAttribute(dynconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmemBytes));
    MCG_CUDA(cudaFuncSetAttribute(dynconv_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynMmaSmemBytes));
    MCG_CUDA(cudaFuncSetAttribute(dynconv_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynMmaSmemBytes));
    MCG_CUDA(cudaFuncSetAttribute(small_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MCG_CUDA(cudaFuncSetAttribute(linear256_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MCG_CUDA(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemBytes));
  }

  ~Engine() {
    drop_graph();
    for (int i = 0; i < 2; ++i) {
      if (slot_[i].h2d_done) cudaEventDestroy(slot_[i].h2d_done);
      if (slot_[i].done) cudaEventDestroy(slot_[i].done);
      if (slot_[i].pin_out) cudaFreeHost(slot_[i].pin_out);
    }
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (side_stream_) {
      cudaStreamDestroy(side_stream_);
      for (int i = 0; i < 4; ++i) {
        cudaEventDestroy(fork_ev_[i]);
        cudaEventDestroy(join_ev_[i]);
      }
      cudaEventDestroy(split_fork_ev_);
      cudaEventDestroy(split_join_ev_);
    }
    for (cudaEvent_t e : ev_pool_) cudaEventDestroy(e);
    for (cudaEvent_t e : prof_ev_) cudaEventDestroy(e);
    if (pin_meta_) cudaFreeHost(pin_meta_);
    for (cudaEvent_t e : meta_ev_)
      if (e) cudaEventDestroy(e);
    if (own_stream_) cudaStreamDestroy(own_stream_);
    if (cap_stream_) cudaStreamDestroy(cap_stream_);
  }
  void set_option(const std::string& k, int v) {
    if (k == "keep_intermediates") keep_stage_interm_ = v != 0;
    else if (k == "time_kernels") time_kernels_ = v != 0;
    else if (k == "head_tensor_cores") { head_tc_ = v != 0; drop_graph(true); }
    else if (k == "fused_stem") { fused_stem_ = v != 0; drop_graph(); }
    else if (k == "fuse_downsample") { fuse_ds_ = v != 0; drop_graph(); }
    else if (k == "fuse_bottleneck") { fuse_bneck_ = v != 0; drop_graph(); }
    else if (k == "split_layers") { split_layers_ = v < 0 ? default_split_layers() : (v & 0xf); drop_graph(); }
    else if (k == "split_min_frames") { split_min_frames_ = v < 0 ? default_split_min_frames() : v; drop_graph(); }
    else throw CudaError("check failed: unknown option " + k);
  }

  void set_graph_mode(int on) {
    graph_mode_ = on != 0;
    drop_graph();
  }
  int last_launches() const { return launches_; }
  // stats of the tcgen05 GEMM launches of the last eager forward: [0] launches, [1] algorithmic
  // FLOPs, [2] summed device time in ms (needs option time_kernels=1; synchronises)
  void umma_stats(double out[3]) {
    out[0] = umma_launches_;
    out[1] = umma_flops_;
    double ms = 0.0;
    if (ev_used_) {
      MCG_CUDA(cudaEventSynchronize(ev_pool_[ev_used_ - 1]));
      for (size_t i = 0; i + 1 < ev_used_; i += 2) {
        float t = 0.f;
        MCG_CUDA(cudaEventElapsedTime(&t, ev_pool_[i], ev_pool_[i + 1]));
        ms += t;
      }
    }
    out[2] = ms;
  }

  void forward(const float* img, int B, int T, int H, int W, const float* img_hw, const float* scale_factor,
               float* out_gaze, float* out_boxes, float* out_scores, cudaStream_t stream) {
    MCG_CUDA(cudaSetDevice(device_));
    MCG_CHECK(B > 0 && T > 0 && H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, "H and W must be multiples of 32");
    const int NB = B * T;
    ensure_workspace(NB, T, H, W);
    // per-call host metadata -> device (tiny)
    std::vector<float> meta(static_cast<size_t>(NB) * 6);
    for (int i = 0; i < NB; ++i) {
      meta[i * 2] = img_hw ? img_hw[i * 2] : static_cast<float>(H);
      meta[i * 2 + 1] = img_hw ? img_hw[i * 2 + 1] : static_cast<float>(W);
      for (int j = 0; j < 4; ++j) meta[NB * 2 + i * 4 + j] = scale_factor ? scale_factor[i * 4 + j] : 1.f;
    }
    if (meta != meta_host_) {
      // metadata changed (every batch of an evaluation run: per-frame crops give per-frame scale factors): stage it in
      // the next slot of a small pinned ring; only the copy issued kMetaSlots changes ago must have completed, so the
      // host does not wait for the forward that is still running
      const int slot = meta_slot_;
      meta_slot_ = (meta_slot_ + 1) % kMetaSlots;
      if (!meta_ev_[slot]) MCG_CUDA(cudaEventCreateWithFlags(&meta_ev_[slot], cudaEventDisableTiming));
      else MCG_CUDA(cudaEventSynchronize(meta_ev_[slot]));
      // slots have ONE stride whatever NB is (the ring's capacity / kMetaSlots): with a per-call stride a short batch's
      // slot overlapped the still-pending copy of the longer batch queued before it (seen once the host stopped waiting
      // for uploads: device-side PNG decode)
      float* pin = pin_meta_ + static_cast<size_t>(slot) * (pin_meta_bytes_ / sizeof(float) / kMetaSlots);
      std::memcpy(pin, meta.data(), meta.size() * sizeof(float));
      MCG_CUDA(cudaMemcpyAsync(d_meta_, pin, meta.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
      MCG_CUDA(cudaEventRecord(meta_ev_[slot], stream));
      meta_host_ = meta;
    }
    has_scale_ = scale_factor != nullptr;

    if (graph_mode_) {
      const GraphKey key{img, out_gaze, out_boxes, out_scores, has_scale_};
      auto it = graphs_.find(key);
      if (it == graphs_.end()) {
        if (graphs_.size() >= 8) drop_graph();
        // capture on a private stream (the caller's may be the legacy default stream, which cannot
        // be captured), then replay the instantiated graph on the caller's stream
        if (!cap_stream_) MCG_CUDA(cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking));
        MCG_CUDA(cudaStreamSynchronize(stream));
        cudaGraph_t graph = nullptr;
        MCG_CUDA(cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal));
        try {
          schedule(img, out_gaze, out_boxes, out_scores, cap_stream_);
        } catch (...) {
          cudaStreamEndCapture(cap_stream_, &graph);
          if (graph) cudaGraphDestroy(graph);
          throw;
        }
        MCG_CUDA(cudaStreamEndCapture(cap_stream_, &graph));
        cudaGraphExec_t exec = nullptr;
        MCG_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        MCG_CUDA(cudaGraphDestroy(graph));
        it = graphs_.emplace(key, exec).first;
      }
      MCG_CUDA(cudaGraphLaunch(it->second, stream));
      return;
    }
    schedule(img, out_gaze, out_boxes, out_scores, stream);
  }

  // Two-slot pipelined host entry: the H2D copy of submission i+1 (copy stream) overlaps the forward
  // of submission i (compute stream); results come back through pinned slot buffers.
  int submit_host(const float* img_host, int B, int T, int H, int W, const float* img_hw, const float* scale_factor) {
    MCG_CUDA(cudaSetDevice(device_));
    const int NB = B * T;
    const size_t in_bytes = static_cast<size_t>(NB) * 3 * H * W * sizeof(float);
    const size_t out_bytes = static_cast<size_t>(NB) * (12 + 12 + 3) * sizeof(float);
    if (!own_stream_) MCG_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    if (!copy_stream_) MCG_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    const int si = next_slot_;
    HostSlot& sl = slot_[si];
    MCG_CHECK(!sl.busy, "mcg_submit_host: both pipeline slots are in flight; call mcg_wait_host first");
    if (in_bytes > sl.in_bytes || out_bytes > sl.out_bytes) {
      MCG_CUDA(cudaDeviceSynchronize());
      drop_graph();
      sl.d_in.reset(new DeviceBlock(in_bytes));
      sl.d_out.reset(new DeviceBlock(out_bytes));
      sl.in_bytes = in_bytes;
      sl.out_bytes = out_bytes;
      if (sl.pin_out) cudaFreeHost(sl.pin_out);
      MCG_CUDA(cudaMallocHost(&sl.pin_out, out_bytes));
    }
    if (!sl.h2d_done) {
      MCG_CUDA(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
      MCG_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    } else {
      MCG_CUDA(cudaStreamWaitEvent(copy_stream_, sl.done, 0));  // previous forward on this slot has consumed d_in
    }
    float* d_in = reinterpret_cast<float*>(sl.d_in->p);
    float* d_out = reinterpret_cast<float*>(sl.d_out->p);
    // asynchronous when the caller's buffer is pinned; otherwise the runtime stages it
    MCG_CUDA(cudaMemcpyAsync(d_in, img_host, in_bytes, cudaMemcpyHostToDevice, copy_stream_));
    MCG_CUDA(cudaEventRecord(sl.h2d_done, copy_stream_));
    MCG_CUDA(cudaStreamWaitEvent(own_stream_, sl.h2d_done, 0));
    forward(d_in, B, T, H, W, img_hw, scale_factor, d_out, d_out + NB * 12, d_out + NB * 24, own_stream_);
    MCG_CUDA(cudaMemcpyAsync(sl.pin_out, d_out, out_bytes, cudaMemcpyDeviceToHost, own_stream_));
    MCG_CUDA(cudaEventRecord(sl.done, own_stream_));
    sl.busy = true;
    sl.NB = NB;
    next_slot_ ^= 1;
    return si;
  }

  void wait_host(int ticket, float* out_gaze, float* out_boxes, float* out_scores) {
    MCG_CHECK(ticket == 0 || ticket == 1, "mcg_wait_host: bad ticket");
    HostSlot& sl = slot_[ticket];
    MCG_CHECK(sl.busy, "mcg_wait_host: nothing submitted on this ticket");
    MCG_CUDA(cudaEventSynchronize(sl.done));
    const int NB = sl.NB;
    const float* po = reinterpret_cast<const float*>(sl.pin_out);
    std::memcpy(out_gaze, po, static_cast<size_t>(NB) * 12 * sizeof(float));
    std::memcpy(out_boxes, po + NB * 12, static_cast<size_t>(NB) * 12 * sizeof(float));
    std::memcpy(out_scores, po + NB * 24, static_cast<size_t>(NB) * 3 * sizeof(float));
    sl.busy = false;
  }

  void forward_host(const float* img_host, int B, int T, int H, int W, const float* img_hw,
                    const float* scale_factor, float* out_gaze, float* out_boxes, float* out_scores) {
    const int t = submit_host(img_host, B, T, H, W, img_hw, scale_factor);
    wait_host(t, out_gaze, out_boxes, out_scores);
  }

  // per-launch device times (ms) of the tcgen05 GEMMs of the last eager forward (option time_kernels)
  int umma_times(double* out, int cap) {
    int n = 0;
    if (ev_used_) MCG_CUDA(cudaEventSynchronize(ev_pool_[ev_used_ - 1]));
    for (size_t i = 0; i + 1 < ev_used_ && n < cap; i += 2, ++n) {
      float t = 0.f;
      MCG_CUDA(cudaEventElapsedTime(&t, ev_pool_[i], ev_pool_[i + 1]));
      out[n] = t;
    }
    return n;
  }

  // "name<TAB>ms" lines, one per kernel launch of the last eager forward (option time_kernels)
  std::string kernel_profile() {
    std::string out;
    if (prof_used_ < 2) return out;
    MCG_CUDA(cudaEventSynchronize(prof_ev_[prof_used_ - 1]));
    for (size_t i = 1; i < prof_used_; ++i) {
      float t = 0.f;
      MCG_CUDA(cudaEventElapsedTime(&t, prof_ev_[i - 1], prof_ev_[i]));
      out += prof_names_[i] + "\t" + std::to_string(t) + "\n";
    }
    return out;
  }

  int get_intermediate(const char* name, float* dst, int64_t capacity, int64_t shape_out[4]) {
    auto it = interm_.find(name);
    if (it == interm_.end()) return MCG_ERR_INVALID;
    const Interm& t = it->second;
    int64_t cnt = 1;
    for (int i = 0; i < 4; ++i)
      if (t.shape[i] > 0) cnt *= t.shape[i];
    if (cnt > capacity) return MCG_ERR_INVALID;
    if (t.kind == 0) {
      MCG_CUDA(cudaMemcpy(dst, t.f32, cnt * sizeof(float), cudaMemcpyDeviceToDevice));
      for (int i = 0; i < 4; ++i) shape_out[i] = t.shape[i];
    } else {
      const int NB = static_cast<int>(t.shape[0]), H = static_cast<int>(t.shape[1]), W = static_cast<int>(t.shape[2]),
                C = static_cast<int>(t.shape[3]);
      planes_to_nchw_kernel<<<1024, 256>>>(t.pl.hi, t.pl.lo, t.pl.lo8, NB, H, W, C, dst);
      MCG_CUDA(cudaGetLastError());
      MCG_CUDA(cudaDeviceSynchronize());
      shape_out[0] = NB;
      shape_out[1] = C;
      shape_out[2] = H;
      shape_out[3] = W;
    }
    return MCG_OK;
  }

  // Value ranges of the trunk / FPN activations of the LAST forward (every tensor has its own arena buffer, so they
  // are all still there) and of the BN-folded convolution weights, as text lines
  // "name<TAB>total<TAB>nonzero<TAB>over<TAB>under<TAB>nonfinite<TAB>maxabs<TAB>energy<TAB>energy_under".
  // Activations: over = |v| > 448, under = 0 < |v| < 2^-8; weights: over = |w| > 28, under = 0 < |w| < 2^-10.
  std::string range_report() {
    MCG_CHECK(ws_NB_ > 0, "mcg_range_report: run a forward first");
    MCG_CUDA(cudaSetDevice(device_));
    MCG_CUDA(cudaDeviceSynchronize());
    std::vector<std::pair<std::string, const Act*>> acts;
    acts.emplace_back("pool", &pool_out_);
    for (int l = 0; l < 4; ++l)
      for (size_t b = 0; b < blk_act_[l].size(); ++b) {
        const std::string k = "layer" + std::to_string(l + 1) + "." + std::to_string(b);
        acts.emplace_back(k + ".t1", &blk_act_[l][b].t1);
        if (!fused_tail_used(l, b)) acts.emplace_back(k + ".t2", &blk_act_[l][b].t2);
        if (blocks_[l][b].has_ds && !(fuse_ds_ && trunk_terms() != 0)) acts.emplace_back(k + ".ds", &blk_act_[l][b].ds);
        acts.emplace_back(k, &blk_act_[l][b].out);
      }
    for (int i = 0; i < 4; ++i) acts.emplace_back("lat" + std::to_string(i), &lat_[i]);
    for (int i = 0; i < 4; ++i) acts.emplace_back("fpn" + std::to_string(i), &fpn_[i]);
    DeviceBlock cnt(acts.size() * 8 * sizeof(unsigned long long));
    MCG_CUDA(cudaMemset(cnt.p, 0, cnt.bytes));
    for (size_t i = 0; i < acts.size(); ++i) {
      unsigned long long* c = reinterpret_cast<unsigned long long*>(cnt.p) + i * 8;
      const long long n = acts[i].second->rows() * acts[i].second->C;
      range_scan_kernel<<<num_sms_ * 4, 256>>>(acts[i].second->pl.hi, n, c, reinterpret_cast<double*>(c + 5));
    }
    MCG_CUDA(cudaGetLastError());
    std::vector<unsigned long long> host(acts.size() * 8);
    MCG_CUDA(cudaMemcpy(host.data(), cnt.p, cnt.bytes, cudaMemcpyDeviceToHost));
    std::vector<RangeRow> rows;
    for (size_t i = 0; i < acts.size(); ++i) {
      RangeRow r;
      r.name = acts[i].first;
      r.total = static_cast<unsigned long long>(acts[i].second->rows() * acts[i].second->C);
      r.nonzero = host[i * 8];
      r.over = host[i * 8 + 1];
      r.under = host[i * 8 + 2];
      r.nonfinite = host[i * 8 + 3];
      const unsigned int mb = static_cast<unsigned int>(host[i * 8 + 4]);
      std::memcpy(&r.maxabs, &mb, 4);
      std::memcpy(&r.energy, &host[i * 8 + 5], 8);
      std::memcpy(&r.energy_under, &host[i * 8 + 6], 8);
      rows.push_back(r);
    }
    rows.insert(rows.end(), weight_ranges_.begin(), weight_ranges_.end());
    std::string out;
    char line[512];
    for (const RangeRow& r : rows) {
      std::snprintf(line, sizeof(line), "%s\t%llu\t%llu\t%llu\t%llu\t%llu\t%.9g\t%.9g\t%.9g\n", r.name.c_str(), r.total, r.nonzero,
                    r.over, r.under, r.nonfinite, static_cast<double>(r.maxabs), r.energy, r.energy_under);
      out += line;
    }
    return out;
  }

 private:
  // whether block (l, b) runs conv2 -> conv3 as one fused kernel (its t2 tensor then stays on chip)
  bool fused_tail_used(int l, size_t b) const {
    const BlockW& bw = blocks_[l][b];
    return fuse_bneck_ && precision_ == MCG_PRECISION_FP16C8 && bneck_supported(bw.c2.Cout, bw.c2.stride, bw.has_ds);
  }

  // -------------------------------------------------------------------------- weight loading
  struct HostT {
    const float* p;
    size_t n;
  };
  const HostT& need(const std::string& k, size_t count) {
    auto it = host_.find(k);
    if (it == host_.end()) throw CudaError("missing checkpoint key: " + k);
    if (it->second.n != count)
      throw CudaError("checkpoint key " + k + " has " + std::to_string(it->second.n) + " elements, expected " +
                      std::to_string(count));
    return it->second;
  }
  bool has(const std::string& k) const { return host_.count(k) != 0; }

  // e4m3 operand planes of the fp16c8 correction MMAs (fixed power-of-two scales, common.cuh)
  void pack_c8(GemmW& g, const std::vector<float>& w, const std::vector<__half>& hi) {
    std::vector<uint8_t> hi8(w.size()), lo8(w.size());
    pack_w8_host(w.data(), hi.data(), w.size(), hi8.data(), lo8.data());
    g.w.hi8 = upload(keep_, hi8);
    g.w.lo8 = upload(keep_, lo8);
  }

  GemmW pack_gemm(const std::vector<float>& w, int N, int K, const float* bias, bool c8 = false) {
    GemmW g;
    g.N = N;
    g.K = K;
    g.w_f32 = upload(keep_, w);
    std::vector<__half> hi(w.size()), lo(w.size());
    for (size_t i = 0; i < w.size(); ++i) {
      const __half h = h_from_float(w[i]);
      hi[i] = h;
      lo[i] = h_from_float(w[i] - __half2float(h));
    }
    g.w.hi = upload(keep_, hi);
    g.w.lo = upload(keep_, lo);
    if (c8) pack_c8(g, w, hi);
    if (bias) g.bias = upload(keep_, std::vector<float>(bias, bias + N));
    if (N <= 768 && K <= 2048 && K % 32 == 0) {
      std::vector<float> t(w.size());
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) t[static_cast<size_t>(k) * N + n] = w[static_cast<size_t>(n) * K + k];
      g.w_t = upload(keep_, t);
    }
    return g;
  }

  // conv weight [Cout,Cin,R,S] (+ optional BN, + optional conv bias) -> [Cout, (r,s,c)] scaled, bias folded
  // stem_order: the 7x7 stem uses the K order of stem_k_index (head_kernels.cuh) instead of (r, s, c)
  ConvW pack_conv(const std::string& wkey, const std::string& bnkey, const std::string& biaskey, int Cout, int Cin,
                  int R, int S, int stride, int pad, int Kpad = 0, bool stem_order = false) {
    const float* w = need(wkey, static_cast<size_t>(Cout) * Cin * R * S).p;
    std::vector<float> scale(Cout, 1.f), shift(Cout, 0.f);
    if (!bnkey.empty()) {
      const float* g = need(bnkey + ".weight", Cout).p;
      const float* b = need(bnkey + ".bias", Cout).p;
      const float* m = need(bnkey + ".running_mean", Cout).p;
      const float* v = need(bnkey + ".running_var", Cout).p;
      for (int o = 0; o < Cout; ++o) {
        // eval-mode BN (eps 1e-5) folded in double, as tools/test.py --fuse-conv-bn sanctions
        const double s = static_cast<double>(g[o]) / std::sqrt(static_cast<double>(v[o]) + 1e-5);
        scale[o] = static_cast<float>(s);
        shift[o] = static_cast<float>(static_cast<double>(b[o]) - static_cast<double>(m[o]) * s);
      }
    }
    if (!biaskey.empty()) {
      const float* cb = need(biaskey, Cout).p;
      for (int o = 0; o < Cout; ++o) shift[o] += cb[o] * scale[o];
    }
    const int K = R * S * Cin;
    const int Kp = Kpad > 0 ? Kpad : K;
    std::vector<float> packed(static_cast<size_t>(Cout) * Kp, 0.f);
    for (int o = 0; o < Cout; ++o)
      for (int c = 0; c < Cin; ++c)
        for (int r = 0; r < R; ++r)
          for (int s = 0; s < S; ++s)
            packed[static_cast<size_t>(o) * Kp + (stem_order ? stem_k_index(r, s, c) : (r * S + s) * Cin + c)] =
                w[((static_cast<size_t>(o) * Cin + c) * R + r) * S + s] * scale[o];
    ConvW cw;
    cw.g = pack_gemm(packed, Cout, Kp, shift.data(), precision_ == MCG_PRECISION_FP16C8);
    if (precision_ == MCG_PRECISION_FP16C8) {
      // e4m3 range of the weight planes (common.cuh): hi8 = e4m3(W 2^4) saturates above 28, is subnormal below 2^-10
      RangeRow r;
      r.name = "w:" + wkey;
      r.total = packed.size();
      for (float v : packed) {
        const float a = std::fabs(v);
        if (!(a <= 65504.f)) {
          ++r.nonfinite;
          continue;
        }
        if (a != 0.f) ++r.nonzero;
        if (a > 28.f) ++r.over;
        if (a != 0.f && a < 0.0009765625f) {
          ++r.under;
          r.energy_under += static_cast<double>(a) * a;
        }
        r.energy += static_cast<double>(a) * a;
        r.maxabs = std::max(r.maxabs, a);
      }
      weight_ranges_.push_back(r);
    }
    cw.Cin = Cin;
    cw.Cout = Cout;
    cw.R = R;
    cw.S = S;
    cw.stride = stride;
    cw.pad = pad;
    return cw;
  }

  // [N, K1] and [N, K2] packed weights -> [N, K1 + K2] (the fp32 copies are on the device: read them back once)
  ConvW concat_k(const ConvW& a, const ConvW& b) {
    MCG_CHECK(a.g.N == b.g.N, "K-concatenation needs equal output channels");
    const int N = a.g.N, K1 = a.g.K, K2 = b.g.K;
    std::vector<float> wa(static_cast<size_t>(N) * K1), wb(static_cast<size_t>(N) * K2), ba(N), bb(N);
    MCG_CUDA(cudaMemcpy(wa.data(), a.g.w_f32, wa.size() * 4, cudaMemcpyDeviceToHost));
    MCG_CUDA(cudaMemcpy(wb.data(), b.g.w_f32, wb.size() * 4, cudaMemcpyDeviceToHost));
    MCG_CUDA(cudaMemcpy(ba.data(), a.g.bias, N * 4, cudaMemcpyDeviceToHost));
    MCG_CUDA(cudaMemcpy(bb.data(), b.g.bias, N * 4, cudaMemcpyDeviceToHost));
    std::vector<float> w(static_cast<size_t>(N) * (K1 + K2));
    for (int n = 0; n < N; ++n) {
      std::memcpy(&w[static_cast<size_t>(n) * (K1 + K2)], &wa[static_cast<size_t>(n) * K1], K1 * sizeof(float));
      std::memcpy(&w[static_cast<size_t>(n) * (K1 + K2) + K1], &wb[static_cast<size_t>(n) * K2], K2 * sizeof(float));
      ba[n] += bb[n];
    }
    ConvW c = a;
    c.g = pack_gemm(w, N, K1 + K2, ba.data(), precision_ == MCG_PRECISION_FP16C8);
    c.Cin = K1 + K2;
    return c;
  }

  GemmW pack_linear(const std::string& wkey, const std::string& bkey, int N, int K) {
    const float* w = need(wkey, static_cast<size_t>(N) * K).p;
    const float* b = bkey.empty() ? nullptr : need(bkey, N).p;
    return pack_gemm(std::vector<float>(w, w + static_cast<size_t>(N) * K), N, K, b);
  }
  LnW pack_ln(const std::string& key, int C) {
    LnW l;
    l.C = C;
    l.g = upload(keep_, std::vector<float>(need(key + ".weight", C).p, need(key + ".weight", C).p + C));
    l.b = upload(keep_, std::vector<float>(need(key + ".bias", C).p, need(key + ".bias", C).p + C));
    return l;
  }

  void load_weights() {
    stem_ = pack_conv("backbone.conv1.weight", "backbone.bn1", "", 64, 3, 7, 7, 2, 3, kStemK, true);
    const int nblk[4] = {3, 4, 6, 3};
    const int planes[4] = {64, 128, 256, 512};
    int cin = 64;
    for (int l = 0; l < 4; ++l) {
      for (int b = 0; b < nblk[l]; ++b) {
        const std::string p = "backbone.layer" + std::to_string(l + 1) + "." + std::to_string(b);
        const int stride = (b == 0 && l > 0) ? 2 : 1;
        BlockW bw;
        bw.c1 = pack_conv(p + ".conv1.weight", p + ".bn1", "", planes[l], cin, 1, 1, 1, 0);
        bw.c2 = pack_conv(p + ".conv2.weight", p + ".bn2", "", planes[l], planes[l], 3, 3, stride, 1);
        bw.c3 = pack_conv(p + ".conv3.weight", p + ".bn3", "", planes[l] * 4, planes[l], 1, 1, 1, 0);
        bw.has_ds = has(p + ".downsample.0.weight");
        if (bw.has_ds)
          bw.ds = pack_conv(p + ".downsample.0.weight", p + ".downsample.1", "", planes[l] * 4, cin, 1, 1, stride, 0);
        if (bw.has_ds) bw.c3ds = concat_k(bw.c3, bw.ds);
        blocks_[l].push_back(bw);
        cin = planes[l] * 4;
      }
    }
    const int fin[4] = {256, 512, 1024, 2048};
    for (int i = 0; i < 4; ++i) {
      const std::string l = "neck.lateral_convs." + std::to_string(i) + ".conv";
      const std::string f = "neck.fpn_convs." + std::to_string(i) + ".conv";
      lateral_[i] = pack_conv(l + ".weight", "", l + ".bias", 256, fin[i], 1, 1, 1, 0);
      fpnconv_[i] = pack_conv(f + ".weight", "", f + ".bias", 256, 256, 3, 3, 1, 1);
    }
    init_boxes_ = upload(keep_, std::vector<float>(need("rpn_head.init_proposal_bboxes.weight", 12).p,
                                                   need("rpn_head.init_proposal_bboxes.weight", 12).p + 12));
    init_feats_ = upload(keep_, std::vector<float>(need("rpn_head.init_proposal_features.weight", 768).p,
                                                   need("rpn_head.init_proposal_features.weight", 768).p + 768));
    const char* clue[3] = {"face", "eyes", "head"};
    for (int s = 0; s < 4; ++s) {
      const std::string p = "roi_head.bbox_head." + std::to_string(s);
      StageW& st = stage_[s];
      st.in_proj = pack_linear(p + ".attention.attn.in_proj_weight", p + ".attention.attn.in_proj_bias", 768, 256);
      st.out_proj = pack_linear(p + ".attention.attn.out_proj.weight", p + ".attention.attn.out_proj.bias", 256, 256);
      st.attn_norm = pack_ln(p + ".attention_norm", 256);
      const std::string q = p + ".instance_interactive_conv";
      if (dyn_mma_) {
        // rows of dynamic_layer permuted so that its output is [PinT (64 n x 256 k) | PoutT (256 n x 64 k)] instead of
        // the reference's [Pin (256 k x 64 n) | Pout (64 k x 256 n)] (transformer.py:1126-1130): both bmm B operands
        // become K-contiguous for the tensor-core DynamicConv kernel
        const float* w = need(q + ".dynamic_layer.weight", static_cast<size_t>(32768) * 256).p;
        const float* b = need(q + ".dynamic_layer.bias", 32768).p;
        std::vector<float> wp(static_cast<size_t>(32768) * 256), bp(32768);
        auto move_row = [&](int dst, int src) {
          std::memcpy(&wp[static_cast<size_t>(dst) * 256], w + static_cast<size_t>(src) * 256, 256 * sizeof(float));
          bp[dst] = b[src];
        };
        for (int k = 0; k < 256; ++k)
          for (int n = 0; n < 64; ++n) move_row(n * 256 + k, k * 64 + n);
        for (int k = 0; k < 64; ++k)
          for (int n = 0; n < 256; ++n) move_row(16384 + n * 64 + k, 16384 + k * 256 + n);
        st.dyn = pack_gemm(wp, 32768, 256, bp.data());
      } else {
        st.dyn = pack_linear(q + ".dynamic_layer.weight", q + ".dynamic_layer.bias", 32768, 256);
      }
      st.norm_in = pack_ln(q + ".norm_in", 64);
      st.norm_out = pack_ln(q + ".norm_out", 256);
      st.fc = pack_linear(q + ".fc_layer.weight", q + ".fc_layer.bias", 256, 12544);
      st.fc_norm = pack_ln(q + ".fc_norm", 256);
      st.iic_norm = pack_ln(p + ".instance_interactive_conv_norm", 256);
      st.ffn1 = pack_linear(p + ".ffn.layers.0.0.weight", p + ".ffn.layers.0.0.bias", 2048, 256);
      st.ffn2 = pack_linear(p + ".ffn.layers.1.weight", p + ".ffn.layers.1.bias", 256, 2048);
      st.ffn_norm = pack_ln(p + ".ffn_norm", 256);
      st.cls_fc = pack_linear(p + ".cls_fcs.0.weight", "", 256, 256);
      st.cls_ln = pack_ln(p + ".cls_fcs.1", 256);
      for (int j = 0; j < 3; ++j) {
        st.reg_fc[j] = pack_linear(p + ".reg_fcs." + std::to_string(3 * j) + ".weight", "", 256, 256);
        st.reg_ln[j] = pack_ln(p + ".reg_fcs." + std::to_string(3 * j + 1), 256);
        st.fc_cls[j] = pack_linear(p + "." + clue[j] + "_fc_cls.weight", p + "." + clue[j] + "_fc_cls.bias", 1, 256);
        st.fc_reg[j] = pack_linear(p + "." + clue[j] + "_fc_reg.weight", p + "." + clue[j] + "_fc_reg.bias", 4, 256);
      }
    }
    {  // only the last stage's gaze head runs at test time (multiclue_gaze_roi_head.py:377-378)
      const std::string h = "roi_head.gaze_head.3";
      for (int c = 0; c < 3; ++c) {
        const std::string t = h + ".gaze_" + clue[c] + "_fcs";
        const std::string ct = h + ".gaze_" + clue[c] + "_confidence";
        for (int j = 0; j < 2; ++j) {
          gaze_.tower[c][j] = pack_linear(t + "." + std::to_string(3 * j) + ".weight", "", 256, 256);
          gaze_.tower_ln[c][j] = pack_ln(t + "." + std::to_string(3 * j + 1), 256);
          gaze_.ctower[c][j] = pack_linear(ct + "." + std::to_string(3 * j) + ".weight", "", 256, 256);
          gaze_.ctower_ln[c][j] = pack_ln(ct + "." + std::to_string(3 * j + 1), 256);
        }
        gaze_.fc[c] = pack_linear(h + ".fc_" + clue[c] + ".weight", h + ".fc_" + clue[c] + ".bias", 3, 256);
        gaze_.fc_conf[c] = pack_linear(h + ".fc_" + clue[c] + "_confidence.weight",
                                       h + ".fc_" + clue[c] + "_confidence.bias", 3, 256);
      }
      gaze_.wg = upload(keep_, std::vector<float>(need(h + ".fc_gaze.weight", 27).p, need(h + ".fc_gaze.weight", 27).p + 27));
      gaze_.bg = upload(keep_, std::vector<float>(need(h + ".fc_gaze.bias", 3).p, need(h + ".fc_gaze.bias", 3).p + 3));
    }
  }

  // -------------------------------------------------------------------------- workspace
  struct Act {  // NHWC planes
    Planes pl;
    int NB = 0, H = 0, W = 0, C = 0;
    long long rows() const { return static_cast<long long>(NB) * H * W; }
  };
  // hi8: the tensor is the input of a convolution in the fp16c8 mode and carries the e4m3 copy of hi
  Act new_act(int NB, int H, int W, int C, bool hi8 = false) {
    Act a;
    a.NB = NB;
    a.H = H;
    a.W = W;
    a.C = C;
    const size_t n = static_cast<size_t>(NB) * H * W * C;
    a.pl = alloc_planes(n, hi8);
    return a;
  }
  // low-part storage of the trunk's activations: fp16 (fp16x3 / simt), e4m3 (fp16c8) or none (fp16)
  bool lo_fp16() const { return precision_ == MCG_PRECISION_FP16X3 || precision_ == MCG_PRECISION_SIMT; }
  bool lo_fp8() const { return precision_ == MCG_PRECISION_FP16C8; }
  Planes alloc_planes(size_t n, bool hi8 = false) {
    Planes pl;
    pl.hi = arena_.alloc<__half>(n);
    pl.lo = lo_fp16() ? arena_.alloc<__half>(n) : nullptr;
    pl.lo8 = lo_fp8() ? arena_.alloc<uint8_t>(n) : nullptr;
    pl.hi8 = (lo_fp8() && hi8) ? arena_.alloc<uint8_t>(n) : nullptr;
    return pl;
  }

  // Shape change (tail batch of an evaluation run, another canvas): the arena only ever grows, so going back and
  // forth between shapes re-uses the memory AND the launch plans / captured graphs of a shape seen before (same
  // layout order -> same addresses).  The device is synchronised only when the arena has to be re-allocated; otherwise
  // everything already queued keeps its pointers into the same allocation and the stream order keeps it safe.
  void ensure_workspace(int NB, int T, int H, int W) {
    if (NB == ws_NB_ && T == ws_T_ && H == ws_H_ && W == ws_W_) return;
    if (ws_NB_ > 0) {
      ShapeCtx& old = ctx_[std::make_tuple(ws_NB_, ws_T_, ws_H_, ws_W_)];
      old.plans.swap(plans_);
      old.bneck_plans.swap(bneck_plans_);
      old.graphs.swap(graphs_);
      old.stem_plan = stem_plan_;
      old.stem_valid = stem_plan_valid_;
    }
    plans_.clear();
    bneck_plans_.clear();
    graphs_.clear();
    stem_plan_valid_ = false;
    interm_.clear();
    arena_.begin_measure();
    layout_workspace(NB, H, W);
    const size_t bytes = arena_.end_measure() + 4096;
    if (bytes > arena_.capacity()) {
      MCG_CUDA(cudaDeviceSynchronize());
      for (auto& kv : ctx_)
        for (auto& g : kv.second.graphs) cudaGraphExecDestroy(g.second);
      ctx_.clear();                      // their plans and graphs point into the allocation that goes away
      dbg_.clear();
      arena_.reserve(bytes);
    } else {
      arena_.reserve(0);                 // rewind
      auto it = ctx_.find(std::make_tuple(NB, T, H, W));
      if (it != ctx_.end()) {
        plans_.swap(it->second.plans);
        bneck_plans_.swap(it->second.bneck_plans);
        graphs_.swap(it->second.graphs);
        stem_plan_ = it->second.stem_plan;
        stem_plan_valid_ = it->second.stem_valid;
        ctx_.erase(it);
      }
    }
    layout_workspace(NB, H, W);
    ws_NB_ = NB;
    ws_T_ = T;
    ws_H_ = H;
    ws_W_ = W;
    const size_t meta_bytes = static_cast<size_t>(kMetaSlots) * NB * 6 * sizeof(float);
    if (meta_bytes > pin_meta_bytes_) {
      if (pin_meta_) {
        MCG_CUDA(cudaDeviceSynchronize());
        cudaFreeHost(pin_meta_);
      }
      MCG_CUDA(cudaMallocHost(&pin_meta_, meta_bytes));
      pin_meta_bytes_ = meta_bytes;
    }
    meta_host_.clear();
  }

  void layout_workspace(int NB, int H, int W) {
    const int P1 = H / 2, Q1 = W / 2, P2 = H / 4, Q2 = W / 4;
    stemA_ = alloc_planes(static_cast<size_t>(NB) * P1 * Q1 * kStemK);
    stem_out_ = new_act(NB, P1, Q1, 64);
    pool_out_ = new_act(NB, P2, Q2, 64, true);
    int h = P2, w = Q2;
    const int planes_c[4] = {64, 128, 256, 512};
    for (int l = 0; l < 4; ++l) {
      blk_act_[l].clear();
      for (size_t b = 0; b < blocks_[l].size(); ++b) {
        const int stride = blocks_[l][b].c2.stride;
        BlkAct ba;
        ba.t1 = new_act(NB, h, w, planes_c[l], true);
        ba.t2 = new_act(NB, h / stride, w / stride, planes_c[l], true);
        if (blocks_[l][b].has_ds) ba.ds = new_act(NB, h / stride, w / stride, planes_c[l] * 4);  // residual only
        ba.out = new_act(NB, h / stride, w / stride, planes_c[l] * 4, true);
        h /= stride;
        w /= stride;
        blk_act_[l].push_back(ba);
      }
    }
    for (int i = 0; i < 4; ++i) {
      lat_[i] = new_act(NB, H / (4 << i), W / (4 << i), 256, true);
      fpn_[i] = new_act(NB, H / (4 << i), W / (4 << i), 256);
    }
    const size_t Rr = static_cast<size_t>(NB) * 3;
    boxes_[0] = arena_.alloc<float>(Rr * 4);
    boxes_[1] = arena_.alloc<float>(Rr * 4);
    obj_[0] = arena_.alloc<float>(Rr * 256);
    obj_[1] = arena_.alloc<float>(Rr * 256);
    qkv_ = arena_.alloc<float>(Rr * 768);
    att_ = arena_.alloc<float>(Rr * 256);
    xa_ = arena_.alloc<float>(Rr * 256);
    xb_ = arena_.alloc<float>(Rr * 256);
    xc_ = arena_.alloc<float>(Rr * 256);
    params_ = arena_.alloc<float>(Rr * 32768);
    roi_ = arena_.alloc<float>(Rr * 12544);
    roih_.hi = arena_.alloc<__half>(Rr * 12544);  // RoIAlign output as planes (tensor-core DynamicConv)
    roih_.lo = arena_.alloc<__half>(Rr * 12544);
    dynf_ = arena_.alloc<float>(Rr * 12544);
    fc_ = arena_.alloc<float>(Rr * 256);
    fcp_ = arena_.alloc<float>(Rr * 256 * kFcSplit);
    ffn_h_ = arena_.alloc<float>(Rr * 2048);
    t256a_ = arena_.alloc<float>(Rr * 256);
    t256b_ = arena_.alloc<float>(Rr * 256);
    cls_logit_ = arena_.alloc<float>(Rr);
    delta_ = arena_.alloc<float>(Rr * 4);
    gz_a_ = arena_.alloc<float>(static_cast<size_t>(NB) * 256 * 6);  // six gaze-head branches side by side
    gz_b_ = arena_.alloc<float>(static_cast<size_t>(NB) * 256 * 6);
    gvec_ = arena_.alloc<float>(static_cast<size_t>(NB) * 9);
    conf_ = arena_.alloc<float>(static_cast<size_t>(NB) * 9);
    d_meta_ = arena_.alloc<float>(static_cast<size_t>(NB) * 6);
    // split-fp16 staging for the head's tensor-core GEMM operands
    hq_.hi = arena_.alloc<__half>(Rr * 256);
    hq_.lo = arena_.alloc<__half>(Rr * 256);
    hobj_.hi = arena_.alloc<__half>(Rr * 256);  // object features entering a stage (operand of the spatial in_proj)
    hobj_.lo = arena_.alloc<__half>(Rr * 256);
    hx1_.hi = arena_.alloc<__half>(Rr * 256);   // output of the spatial attention (operand of the temporal in_proj)
    hx1_.lo = arena_.alloc<__half>(Rr * 256);
    ffn_p_ = arena_.alloc<float>(Rr * 256 * kFfnSplit);
    hh_.hi = arena_.alloc<__half>(Rr * 2048);
    hh_.lo = arena_.alloc<__half>(Rr * 2048);
    hf_.hi = arena_.alloc<__half>(Rr * 12544);
    hf_.lo = arena_.alloc<__half>(Rr * 12544);
  }

  void ensure_side_stream() {
    if (side_stream_) return;
    MCG_CUDA(cudaStreamCreateWithFlags(&side_stream_, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) {
      MCG_CUDA(cudaEventCreateWithFlags(&fork_ev_[i], cudaEventDisableTiming));
      MCG_CUDA(cudaEventCreateWithFlags(&join_ev_[i], cudaEventDisableTiming));
    }
    MCG_CUDA(cudaEventCreateWithFlags(&split_fork_ev_, cudaEventDisableTiming));
    MCG_CUDA(cudaEventCreateWithFlags(&split_join_ev_, cudaEventDisableTiming));
  }

  // forget every captured graph and (plans = true) every launch plan, of the current shape and of the stashed ones
  void drop_graph(bool plans = false) {
    for (auto& kv : ctx_) {
      for (auto& g : kv.second.graphs) cudaGraphExecDestroy(g.second);
      kv.second.graphs.clear();
      if (plans) {
        kv.second.plans.clear();
        kv.second.bneck_plans.clear();
        kv.second.stem_valid = false;
      }
    }
    if (plans) {
      plans_.clear();
      bneck_plans_.clear();
      stem_plan_valid_ = false;
    }
    for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
    graphs_.clear();
  }

  // -------------------------------------------------------------------------- op helpers
  // every kernel launch of the forward passes through here; with option "time_kernels" (eager mode) an event
  // is recorded after it, so consecutive events bracket each kernel on the (serial) stream
  void count(const char* name) {
    ++launches_;
    if (time_kernels_ && !graph_mode_) {
      if (prof_used_ >= prof_ev_.size()) {
        prof_ev_.push_back(nullptr);
        MCG_CUDA(cudaEventCreate(&prof_ev_.back()));
      }
      MCG_CUDA(cudaEventRecord(prof_ev_[prof_used_++], cur_stream_));
      prof_names_.push_back(name);
    }
  }

  // generic GEMM dispatch.  A: planes (kind 0/1) or fp32 (kind 0).
  // terms: 0 = CUDA-core fp32 kernel, 1 / 3 = tcgen05 kernel with 1 / 3 MMAs per k-step
  int trunk_terms() const {
    switch (precision_) {
      case MCG_PRECISION_SIMT: return 0;
      case MCG_PRECISION_FP16X3: return 3;
      case MCG_PRECISION_FP16C8: return 2;
      default: return 1;
    }
  }
  void gemm(const std::string& key, const Planes* A, const float* A_f32, const AGeom& geom, const GemmW& w,
            long long M, const Epilogue& ep, cudaStream_t st, int terms, int k_split = 1, long long split_stride = 0,
            const Planes* A2 = nullptr, const AGeom* geom2 = nullptr) {
    MCG_CHECK(A2 == nullptr || terms != 0, "the K-concatenated form exists on the tcgen05 path only");
    const bool tensor = terms != 0 && A != nullptr && umma_supported(M, w.N, w.K, geom) &&
                        (terms == 1 || (terms == 3 && A->lo != nullptr) || (terms == 2 && A->lo8 != nullptr && A->hi8 != nullptr && w.w.hi8 != nullptr));
    if (tensor) {
      auto it = plans_.find(key);
      if (it == plans_.end()) {
        // CTA pairs (tcgen05 cta_group::2, 256-row tiles) where the W tile is worth sharing: env MCG_TUNE_PAIR
        // 0 off, 1 the 3x3 / strided convolutions, 2 every trunk convolution with >= 1 tile per pair
        // (measured, profiles/: fp16c8 gains on both, fp16x3 / fp16 only on the 3x3 layers)
        static const int tune_pair = std::getenv("MCG_TUNE_PAIR") ? std::atoi(std::getenv("MCG_TUNE_PAIR")) : -1;
        const int pair_mode = tune_pair >= 0 ? tune_pair : (terms == 2 ? 2 : 1);
        // ... for layers with at least `pair_min` 256-row tiles (env MCG_TUNE_PAIR_MIN_MTILES).  Measured (round 2, same
        // box A/B): 74 (one per SM pair) 9.90 ms per step, 40 (layer4 at 32 clips: 43 tiles) 9.82, 20 9.89
        static const int tune_pair_min = std::getenv("MCG_TUNE_PAIR_MIN_MTILES") ? std::atoi(std::getenv("MCG_TUNE_PAIR_MIN_MTILES")) : 0;
        const long long pair_min = tune_pair_min > 0 ? tune_pair_min : (num_sms_ * 40) / 148;
        const bool big = M >= 2 * kBlockM * pair_min && k_split == 1 && ep.out_f32 == nullptr;
        const int pair = (big && ((pair_mode == 1 && geom.kind == 1) || pair_mode == 2)) ? 1 : 0;
        UmmaPlan pl = make_umma_plan(terms, *A, geom, w.w, M, w.N, w.K, ep, num_sms_, 0, k_split, split_stride, pair, A2, geom2);
        pl.p.reverse = reverse_override_ >= 0 ? reverse_override_ : (alternate_ ? ((tc_launch_index_ & 1) ^ 1) : 0);   // launch 0 reads what the stem wrote last
        it = plans_.emplace(key, pl).first;
      }
      const bool timed = time_kernels_ && !graph_mode_;
      if (timed) {
        if (ev_used_ + 2 > ev_pool_.size()) {
          ev_pool_.resize(ev_used_ + 2);
          MCG_CUDA(cudaEventCreate(&ev_pool_[ev_used_]));
          MCG_CUDA(cudaEventCreate(&ev_pool_[ev_used_ + 1]));
        }
        MCG_CUDA(cudaEventRecord(ev_pool_[ev_used_], st));
      }
      launch_umma(it->second, st);
      ++tc_launch_index_;
      if (timed) {
        MCG_CUDA(cudaEventRecord(ev_pool_[ev_used_ + 1], st));
        ev_used_ += 2;
      }
      ++umma_launches_;
      // algorithmic work: the stem's K is zero-padded from 147 to 192
      umma_flops_ += 2.0 * static_cast<double>(M) * w.N * (key == "stem" ? 147 : w.K);
    } else {
      MCG_CHECK(A2 == nullptr, "K-concatenated GEMM shape not supported by the tcgen05 kernel");
      SimtParams p;
      p.M = M;
      p.N = w.N;
      p.K = w.K;
      p.a = geom;
      if (A) {
        p.a_hi = A->hi;
        p.a_lo = A->lo;
        p.a_lo8 = A->lo8;
      } else {
        p.a_f32 = A_f32;
      }
      p.w_f32 = w.w_f32;
      p.ep = ep;
      launch_simt_gemm(p, st);
    }
    count(tensor ? ("umma:" + key).c_str() : "simt_gemm_kernel");
  }

  // convolution over NHWC planes -> NHWC planes
  // x2 / stride2: second input of a K-concatenated 1x1 convolution pair (conv3 + downsample branch), read with
  // `stride2` on the same output grid
  void conv(const std::string& key, const Act& x, const ConvW& cw, const Act& y, bool relu, const Act* res,
            int res_mode, cudaStream_t st, const Act* x2 = nullptr, int stride2 = 1) {
    AGeom g;
    const bool plain = cw.R == 1 && cw.S == 1 && cw.stride == 1 && cw.pad == 0;
    g.kind = plain ? 0 : 1;
    g.lda = x.C;
    g.NB = x.NB;
    g.H = x.H;
    g.W = x.W;
    g.C = x.C;
    g.R = cw.R;
    g.S = cw.S;
    g.stride = cw.stride;
    g.pad = cw.pad;
    g.P = y.H;
    g.Q = y.W;
    Epilogue ep;
    ep.bias = cw.g.bias;
    ep.relu = relu ? 1 : 0;
    ep.out_hi = y.pl.hi;
    ep.out_lo = y.pl.lo;
    ep.out_lo8 = y.pl.lo8;
    ep.out_hi8 = y.pl.hi8;
    ep.ldo = y.C;
    if (res) {
      ep.res_hi = res->pl.hi;
      ep.res_lo = res->pl.lo;
      ep.res_lo8 = res->pl.lo8;
      ep.res_mode = res_mode;
      ep.ldr = res->C;
      ep.P = y.H;
      ep.Q = y.W;
    }
    if (x2) {
      AGeom g2;
      g2.kind = stride2 == 1 ? 0 : 1;
      g2.lda = x2->C;
      g2.NB = x2->NB;
      g2.H = x2->H;
      g2.W = x2->W;
      g2.C = x2->C;
      g2.stride = stride2;
      g2.P = y.H;
      g2.Q = y.W;
      g.C = x.C;
      gemm(key, &x.pl, nullptr, g, cw.g, y.rows(), ep, st, trunk_terms(), 1, 0, &x2->pl, &g2);
      return;
    }
    gemm(key, &x.pl, nullptr, g, cw.g, y.rows(), ep, st, trunk_terms());
  }

  // frames [f0, f0 + nf) of an NHWC tensor
  static Act slice_frames(const Act& a, int f0, int nf) {
    Act s = a;
    const size_t off = static_cast<size_t>(f0) * a.H * a.W * a.C;
    s.NB = nf;
    s.pl.hi = a.pl.hi + off;
    if (a.pl.lo) s.pl.lo = a.pl.lo + off;
    if (a.pl.lo8) s.pl.lo8 = a.pl.lo8 + off;
    if (a.pl.hi8) s.pl.hi8 = a.pl.hi8 + off;
    return s;
  }

  // One bottleneck (resnet.py:263-302) as two independent chains over the two halves of the frames, chain 0 on `s0`, chain 1
  // on `s1` (see split_layers_).  `chain` numbers the convolutions of a chain for the alternating tile order; -> next number.
  int block_two_chains(int l, size_t b, const Act& x, cudaStream_t s0, cudaStream_t s1, int chain) {
    const BlockW& bw = blocks_[l][b];
    BlkAct& ba = blk_act_[l][b];
    const std::string k = "l" + std::to_string(l) + "b" + std::to_string(b);
    const int n0 = x.NB / 2;
    int c = chain;
    for (int h = 0; h < 2; ++h) {
      cudaStream_t s = h ? s1 : s0;
      const int f0 = h ? n0 : 0, nf = h ? x.NB - n0 : n0;
      const std::string tag = h ? "#1" : "#0";
      const Act xs = slice_frames(x, f0, nf), t1 = slice_frames(ba.t1, f0, nf), t2 = slice_frames(ba.t2, f0, nf),
                out = slice_frames(ba.out, f0, nf);
      c = chain;
      auto order = [&]() { reverse_override_ = alternate_ ? ((c++ & 1) ^ 1) : 0; };
      order();
      conv(k + "c1" + tag, xs, bw.c1, t1, true, nullptr, RES_NONE, s);
      order();
      conv(k + "c2" + tag, t1, bw.c2, t2, true, nullptr, RES_NONE, s);
      if (bw.has_ds && fuse_ds_) {
        order();
        conv(k + "c3ds" + tag, t2, bw.c3ds, out, true, nullptr, RES_NONE, s, &xs, bw.ds.stride);
      } else {
        Act ds;
        const Act* idn = &xs;
        if (bw.has_ds) {
          ds = slice_frames(ba.ds, f0, nf);
          order();
          conv(k + "ds" + tag, xs, bw.ds, ds, false, nullptr, RES_NONE, s);
          idn = &ds;
        }
        order();
        conv(k + "c3" + tag, t2, bw.c3, out, true, idn, RES_SAME, s);
      }
    }
    reverse_override_ = -1;
    return c;
  }

  // fused bottleneck tail: y = relu(conv3(relu(conv2(t1))) + x)
  void bneck(const std::string& key, const Act& t1, const BlockW& bw, const Act& idn, const Act& y, cudaStream_t st) {
    auto it = bneck_plans_.find(key);
    if (it == bneck_plans_.end()) {
      static const int tune_pair = std::getenv("MCG_TUNE_BF_PAIR") ? std::atoi(std::getenv("MCG_TUNE_BF_PAIR")) : 1;
      const long long M = y.rows();
      const int pair = (tune_pair && M >= 2 * kBlockM * ((num_sms_ * 40) / 148)) ? 1 : 0;
      BneckPlan pl = make_bneck_plan(t1.pl, t1.NB, t1.H, t1.W, bw.c2.Cout, bw.c2.g.w, bw.c2.g.bias, bw.c3.g.w, bw.c3.g.bias,
                                     idn.pl, y.pl, num_sms_, pair);
      pl.p.reverse = alternate_ ? ((tc_launch_index_ & 1) ^ 1) : 0;
      it = bneck_plans_.emplace(key, pl).first;
    }
    const bool timed = time_kernels_ && !graph_mode_;
    if (timed) {
      if (ev_used_ + 2 > ev_pool_.size()) {
        ev_pool_.resize(ev_used_ + 2);
        MCG_CUDA(cudaEventCreate(&ev_pool_[ev_used_]));
        MCG_CUDA(cudaEventCreate(&ev_pool_[ev_used_ + 1]));
      }
      MCG_CUDA(cudaEventRecord(ev_pool_[ev_used_], st));
    }
    launch_bneck(it->second, st);
    ++tc_launch_index_;
    if (timed) {
      MCG_CUDA(cudaEventRecord(ev_pool_[ev_used_ + 1], st));
      ev_used_ += 2;
    }
    ++umma_launches_;
    umma_flops_ += it->second.flops;
    count(("bneck:" + key).c_str());
  }

  // fp32 linear on (possibly strided) rows: y = x W^T + b (+res) (relu)
  void linear(const float* x, long long ldx, const GemmW& w, long long M, float* y, long long ldy, bool relu,
              const float* res, long long ldres, cudaStream_t st) {
    if (w.w_t != nullptr) {
      const GemmW* ws[1] = {&w};
      const float* xs[1] = {x};
      float* ys[1] = {y};
      linear_grouped(1, xs, ldx, ws, M, ys, ldy, relu, res, ldres, st);
      return;
    }
    AGeom g;
    g.kind = 0;
    g.lda = ldx;
    Epilogue ep;
    ep.bias = w.bias;
    ep.relu = relu ? 1 : 0;
    ep.out_f32 = y;
    ep.ldo = ldy;
    if (res) {
      ep.res_f32 = res;
      ep.res_mode = RES_SAME;
      ep.ldr = ldres;
    }
    gemm("", nullptr, x, g, w, M, ep, st, 0);
  }

  // n (<= 6) independent small Linears of one shape in a single launch (per-clue heads, gaze branches)
  void linear_grouped(int n, const float* const* x, long long ldx, const GemmW* const* w, long long M, float* const* y,
                      long long ldy, bool relu, const float* res, long long ldres, cudaStream_t st) {
    MCG_CHECK(n >= 1 && n <= kMaxLinGroups && (res == nullptr || n == 1), "bad Linear group");
    LinGroups g = {};
    for (int i = 0; i < n; ++i) {
      MCG_CHECK(w[i]->w_t != nullptr && w[i]->N == w[0]->N && w[i]->K == w[0]->K, "grouped Linears must share a shape");
      g.x[i] = x[i];
      g.wt[i] = w[i]->w_t;
      g.bias[i] = w[i]->bias;
      g.y[i] = y[i];
    }
    const int N = w[0]->N, K = w[0]->K;
    dim3 grid(static_cast<unsigned>((M + kSlRows - 1) / kSlRows), static_cast<unsigned>((N + 63) / 64), n);
    const size_t smem = (static_cast<size_t>(kSlRows) * K + kSlSlices * kSlRows * 64) * sizeof(float);
    small_linear_kernel<<<grid, kSlThreads, smem, st>>>(g, ldx, res, ldres, ldy, M, N, K, relu ? 1 : 0);
    MCG_CUDA(cudaGetLastError());
    count("small_linear_kernel");
  }

  // y = act(LN(x W^T + b (+res)))  for N == 256 Linears followed by a LayerNorm; optionally also written as
  // split-fp16 planes (input of a following tensor-core Linear)
  // attn_qkv != null: x is computed in the kernel as the attention core over qkv (mode 0 spatial / 1 temporal)
  void linear_ln(const float* x, long long ldx, const GemmW& w, const LnW& n, long long M, float* y, long long ldy,
                 bool relu, const float* res, long long ldres, cudaStream_t st, const Planes* planes = nullptr,
                 const float* attn_qkv = nullptr, int attn_T = 0, int attn_mode = 0) {
    const GemmW* ws[1] = {&w};
    const LnW* ns[1] = {&n};
    const float* xs[1] = {x};
    float* ys[1] = {y};
    linear_ln_grouped(1, xs, ldx, ws, ns, M, ys, ldy, relu, res, ldres, st, planes, attn_qkv, attn_T, attn_mode);
  }
  void linear_ln_grouped(int cnt, const float* const* x, long long ldx, const GemmW* const* w, const LnW* const* n,
                         long long M, float* const* y, long long ldy, bool relu, const float* res, long long ldres,
                         cudaStream_t st, const Planes* planes = nullptr, const float* attn_qkv = nullptr, int attn_T = 0,
                         int attn_mode = 0) {
    MCG_CHECK(attn_qkv == nullptr || (cnt == 1 && w[0]->K == 256), "the fused attention core feeds one 256 -> 256 Linear");
    MCG_CHECK(cnt >= 1 && cnt <= kMaxLinGroups && ((res == nullptr && planes == nullptr) || cnt == 1), "bad Linear+LN group");
    LinGroups g = {};
    for (int i = 0; i < cnt; ++i) {
      MCG_CHECK(w[i]->N == 256 && w[i]->w_t != nullptr && n[i]->C == 256 && w[i]->K == w[0]->K,
                "linear_ln needs 256-wide small Linears of one shape");
      g.x[i] = x[i];
      g.wt[i] = w[i]->w_t;
      g.bias[i] = w[i]->bias;
      g.gamma[i] = n[i]->g;
      g.beta[i] = n[i]->b;
      g.y[i] = y[i];
    }
    const int K = w[0]->K;
    const size_t smem = (static_cast<size_t>(kSlRows) * K + 4 * kSlRows * 256) * sizeof(float);
    dim3 grid(static_cast<unsigned>((M + kSlRows - 1) / kSlRows), cnt);
    linear256_ln_kernel<<<grid, 1024, smem, st>>>(g, ldx, res, ldres, ldy, M, K, relu ? 1 : 0,
                                                  planes ? planes->hi : nullptr, planes ? planes->lo : nullptr, attn_qkv,
                                                  attn_T, attn_mode);
    MCG_CUDA(cudaGetLastError());
    count("linear256_ln_kernel");
  }

  // big head linears on tensor cores.  The producer of x normally writes the split-fp16 planes `stage`
  // itself (staged == true); otherwise the fp32 activations are split here.  With out_planes the result
  // goes out as planes (input of the next tensor-core Linear) instead of fp32.
  void linear_tc(const std::string& key, const float* x, int K, const Planes& stage, bool staged, const GemmW& w,
                 long long M, float* y, bool relu, const float* res, cudaStream_t st, int k_split = 1,
                 const Planes* out_planes = nullptr) {
    if (!head_on_tc()) {
      linear(x, K, w, M, y, w.N, relu, res, w.N, st);
      return;
    }
    if (!staged) {
      split_planes_kernel<<<num_sms_ * 4, 256, 0, st>>>(x, K, M, K, stage.hi, stage.lo);
      MCG_CUDA(cudaGetLastError());
      count("split_planes_kernel");
    }
    AGeom g;
    g.kind = 0;
    g.lda = K;
    Epilogue ep;
    ep.bias = w.bias;
    ep.relu = relu ? 1 : 0;
    if (out_planes) {
      ep.out_hi = out_planes->hi;
      ep.out_lo = out_planes->lo;
    } else {
      ep.out_f32 = y;
    }
    ep.ldo = w.N;
    if (res) {
      ep.res_f32 = res;
      ep.res_mode = RES_SAME;
      ep.ldr = w.N;
    }
    if (k_split > 1) ep.bias = nullptr;  // the split-K reduction (fused into the following LayerNorm) adds it
    // head GEMMs always use the 3-term split: the head is precision critical and only 2.5 % of the FLOPs
    gemm(key, &stage, nullptr, g, w, M, ep, st, 3, k_split, static_cast<long long>(M) * w.N);
  }
  bool head_on_tc() const { return precision_ != MCG_PRECISION_SIMT && head_tc_; }

  // second stage (w2 != null): y = LN_w2(res2 + act(LN_w(...)))
  void ln(const float* x, long long ldx, const float* res, long long ldres, const LnW& w, float* y, long long ldy,
          long long rows, bool relu, cudaStream_t st, int nsplit = 1, long long split_stride = 0,
          const float* xbias = nullptr, const Planes* planes = nullptr, const float* res2 = nullptr, long long ldres2 = 0,
          const LnW* w2 = nullptr) {
    const int wpb = 4;  // warps (rows) per block: 168 blocks for 672 rows
    const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
    MCG_CHECK((w.C == 256 || w.C == 64) && (w2 == nullptr || w2->C == w.C), "LayerNorm width must be 64 or 256");
    auto launch = [&](auto kernel) {
      kernel<<<grid, wpb * 32, 0, st>>>(x, ldx, res, ldres, w.g, w.b, y, ldy, rows, relu ? 1 : 0, nsplit, split_stride, xbias,
                                       planes ? planes->hi : nullptr, planes ? planes->lo : nullptr, res2, ldres2,
                                       w2 ? w2->g : nullptr, w2 ? w2->b : nullptr);
    };
    if (w.C == 256)
      launch(layernorm_kernel<8>);
    else
      launch(layernorm_kernel<2>);
    MCG_CUDA(cudaGetLastError());
    count("layernorm_kernel");
  }

  const float* snapshot(const float* src, size_t n, cudaStream_t st) {
    dbg_.emplace_back(new DeviceBlock(n * sizeof(float)));
    MCG_CUDA(cudaMemcpyAsync(dbg_.back()->p, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return reinterpret_cast<const float*>(dbg_.back()->p);
  }
  void reg_f32(const std::string& name, const float* p, int64_t a, int64_t b = 0, int64_t c = 0, int64_t d = 0) {
    Interm t;
    t.kind = 0;
    t.f32 = p;
    t.shape[0] = a;
    t.shape[1] = b;
    t.shape[2] = c;
    t.shape[3] = d;
    interm_[name] = t;
  }
  void reg_act(const std::string& name, const Act& a) {
    Interm t;
    t.kind = 1;
    t.pl = a.pl;
    t.shape[0] = a.NB;
    t.shape[1] = a.H;
    t.shape[2] = a.W;
    t.shape[3] = a.C;
    interm_[name] = t;
  }

  // -------------------------------------------------------------------------- the forward schedule
  void schedule(const float* img, float* out_gaze, float* out_boxes, float* out_scores, cudaStream_t st) {
    launches_ = 0;
    tc_launch_index_ = 0;
    umma_launches_ = 0;
    umma_flops_ = 0.0;
    ev_used_ = 0;
    cur_stream_ = st;
    prof_used_ = 0;
    prof_names_.clear();
    count("(start)");
    --launches_;
    if (!graph_mode_) dbg_.clear();
    const int NB = ws_NB_, T = ws_T_, H = ws_H_, W = ws_W_;
    const int ew_grid = num_sms_ * 8;
    // ---- stem (resnet.py:636-639): one fused tcgen05 kernel (conv 7x7/2 + BN + ReLU + max-pool), or the
    // unfused im2col -> GEMM -> max-pool chain (CUDA-core mode, "fused_stem" = 0: keeps the 112^2 stem map
    // addressable for per-layer parity tests)
    const bool fused_stem = fused_stem_ && precision_ != MCG_PRECISION_SIMT && stem_fused_supported(H, W);
    if (fused_stem) {
      if (!stem_plan_valid_) {
        // fp16c8: the stem (2 % of the FLOPs, builder-bound) keeps the 3-term fp16 products and writes hi + lo8
        stem_plan_ = make_stem_fused_plan(trunk_terms() == 1 ? 1 : 3, img, NB, H, W, stem_.g.w, stem_.g.bias, pool_out_.pl,
                                          num_sms_);
        stem_plan_valid_ = true;
      }
      stem_plan_.p.img = img;
      launch_stem_fused(stem_plan_, st);
      count("stem_fused_kernel");
      // algorithmic work of the stem convolution (K = 147), as the GEMM path counts it
      umma_flops_ += 2.0 * static_cast<double>(stem_out_.rows()) * 64 * 147;
    } else {
      stem_im2col_kernel<<<NB * (H / 2), 256, 3 * 7 * (W + 6) * sizeof(float), st>>>(img, NB, H, W, H / 2, W / 2,
                                                                                     stemA_.hi, stemA_.lo, stemA_.lo8);
      MCG_CUDA(cudaGetLastError());
      count("stem_im2col_kernel");
      {
        AGeom g;
        g.kind = 0;
        g.lda = kStemK;
        Epilogue ep;
        ep.bias = stem_.g.bias;
        ep.relu = 1;
        ep.out_hi = stem_out_.pl.hi;
        ep.out_lo = stem_out_.pl.lo;
        ep.out_lo8 = stem_out_.pl.lo8;
        ep.ldo = 64;
        gemm("stem", &stemA_, nullptr, g, stem_.g, stem_out_.rows(), ep, st, trunk_terms());
      }
      maxpool3x3s2_kernel<<<ew_grid, 256, 0, st>>>(stem_out_.pl.hi, stem_out_.pl.lo, stem_out_.pl.lo8, NB, H / 2, W / 2, 64,
                                                   H / 4, W / 4, pool_out_.pl.hi, pool_out_.pl.lo, pool_out_.pl.lo8, pool_out_.pl.hi8);
      MCG_CUDA(cudaGetLastError());
      count("maxpool3x3s2_kernel");
      reg_act("stem", stem_out_);
    }
    reg_act("pool", pool_out_);
    // ---- layer1..4 (resnet.py:263-302)
    const Act* x = &pool_out_;
    // two-chain region (split_layers_): per-kernel timing mode keeps everything on one stream
    const bool may_split = split_layers_ != 0 && NB >= split_min_frames_ && NB >= 2 && trunk_terms() != 0 &&
                           !(time_kernels_ && !graph_mode_);
    bool forked = false;
    int chain = 0;
    auto join_chains = [&]() {
      if (!forked) return;
      MCG_CUDA(cudaEventRecord(split_join_ev_, side_stream_));
      MCG_CUDA(cudaStreamWaitEvent(st, split_join_ev_, 0));
      reverse_override_ = -1;
      forked = false;
    };
    for (int l = 0; l < 4; ++l) {
      for (size_t b = 0; b < blocks_[l].size(); ++b) {
        const BlockW& bw = blocks_[l][b];
        BlkAct& ba = blk_act_[l][b];
        const std::string k = "l" + std::to_string(l) + "b" + std::to_string(b);
        if (may_split && ((split_layers_ >> l) & 1) && !fused_tail_used(l, b)) {
          if (!forked) {
            ensure_side_stream();
            MCG_CUDA(cudaEventRecord(split_fork_ev_, st));
            MCG_CUDA(cudaStreamWaitEvent(side_stream_, split_fork_ev_, 0));
            forked = true;
            chain = tc_launch_index_;
          }
          chain = block_two_chains(l, b, *x, st, side_stream_, chain);
          x = &ba.out;
          reg_act("layer" + std::to_string(l + 1) + "." + std::to_string(b), ba.out);
          continue;
        }
        join_chains();
        conv(k + "c1", *x, bw.c1, ba.t1, true, nullptr, RES_NONE, st);
        if (fused_tail_used(l, b)) {
          // conv2 -> conv3 + identity as one kernel, t2 stays in shared memory (bneck_fused.cuh)
          bneck(k + "c2c3", ba.t1, bw, *x, ba.out, st);
          x = &ba.out;
          reg_act("layer" + std::to_string(l + 1) + "." + std::to_string(b), ba.out);
          continue;
        }
        conv(k + "c2", ba.t1, bw.c2, ba.t2, true, nullptr, RES_NONE, st);
        const Act* idn = x;
        if (bw.has_ds && fuse_ds_ && trunk_terms() != 0) {
          // conv3 and the downsample branch as ONE GEMM over the concatenated K (t2 channels, then x channels read with
          // the block's stride): out = relu(W3 t2 + Wds x + b3 + bds) (resnet.py:286-295)
          conv(k + "c3ds", ba.t2, bw.c3ds, ba.out, true, nullptr, RES_NONE, st, x, bw.ds.stride);
        } else {
          if (bw.has_ds) {
            conv(k + "ds", *x, bw.ds, ba.ds, false, nullptr, RES_NONE, st);
            idn = &ba.ds;
          }
          conv(k + "c3", ba.t2, bw.c3, ba.out, true, idn, RES_SAME, st);
        }
        x = &ba.out;
        reg_act("layer" + std::to_string(l + 1) + "." + std::to_string(b), ba.out);
      }
    }
    join_chains();
    // ---- FPN (fpn.py:151-180): lateral 1x1 (+ top-down nearest-2x add fused), then 3x3
    for (int i = 3; i >= 0; --i) {
      const Act& c = blk_act_[i].back().out;
      conv("lat" + std::to_string(i), c, lateral_[i], lat_[i], false, i < 3 ? &lat_[i + 1] : nullptr,
           i < 3 ? RES_UP2X : RES_NONE, st);
    }
    for (int i = 0; i < 4; ++i) {
      conv("fpn" + std::to_string(i), lat_[i], fpnconv_[i], fpn_[i], false, nullptr, RES_NONE, st);
      reg_act("fpn" + std::to_string(i), fpn_[i]);
    }
    // ---- query head (multiclue_gaze_roi_head.py:287-384)
    const int R = NB * 3;
    float* img_hw = d_meta_;
    float* scale = d_meta_ + NB * 2;
    const bool tc = head_on_tc();
    init_proposals_kernel<<<NB, 256, 0, st>>>(init_boxes_, init_feats_, img_hw, NB, boxes_[0], obj_[0], tc ? hobj_.hi : nullptr,
                                              tc ? hobj_.lo : nullptr);
    MCG_CUDA(cudaGetLastError());
    count("init_proposals_kernel");
    FpnLevels fl;
    for (int i = 0; i < 4; ++i) {
      fl.hi[i] = fpn_[i].pl.hi;
      fl.lo[i] = fpn_[i].pl.lo;
      fl.lo8[i] = fpn_[i].pl.lo8;
      fl.H[i] = fpn_[i].H;
      fl.W[i] = fpn_[i].W;
    }
    int cur = 0;
    for (int s = 0; s < 4; ++s) {
      const StageW& sw = stage_[s];
      const std::string sk = "s" + std::to_string(s);
      float* boxes_in = boxes_[cur];
      float* boxes_out = boxes_[cur ^ 1];
      float* obj_in = obj_[cur];
      float* obj_out = obj_[cur ^ 1];
      // RoIAlign (needs this stage's boxes) and the attention block (needs the object features) are independent
      // until DynamicConv: RoIAlign runs on a forked side stream (both are latency-bound, neither fills the GPU).
      // Per-kernel timing mode keeps everything on one stream.
      const bool fork = !(time_kernels_ && !graph_mode_);
      cudaStream_t rs = st;
      if (fork) {
        ensure_side_stream();
        MCG_CUDA(cudaEventRecord(fork_ev_[s], st));
        MCG_CUDA(cudaStreamWaitEvent(side_stream_, fork_ev_[s], 0));
        rs = side_stream_;
      }
      roi_align_kernel<<<(R * 49 + 7) / 8, 256, 0, rs>>>(fl, boxes_in, R, roi_, dyn_mma_ ? roih_.hi : nullptr,
                                                          dyn_mma_ ? roih_.lo : nullptr);
      MCG_CUDA(cudaGetLastError());
      count("roi_align_kernel");
      if (fork) MCG_CUDA(cudaEventRecord(join_ev_[s], side_stream_));
      // spatial then temporal self-attention with the SAME weights (gaze_stqi_head.py:148-166)
      const float* xin = obj_in;
      float* xout[2] = {xa_, xb_};
      for (int mode = 0; mode < 2; ++mode) {
        // in_proj (M = 3 frames rows, N = 768, K = 256) on the tcgen05 GEMM: its operand planes come from the kernel
        // that produced xin (init proposals / the previous stage's ffn_norm / the spatial pass's out_proj + LN)
        linear_tc(sk + (mode == 0 ? "inproj_s" : "inproj_t"), xin, 256, mode == 0 ? hobj_ : hx1_, tc, sw.in_proj, R, qkv_,
                  false, nullptr, st);
        // attention core + out_proj + identity (mmcv MHA) + attention_norm in one kernel; it also emits the split-fp16
        // planes the next tensor-core Linear reads (temporal in_proj resp. dynamic_layer)
        linear_ln(nullptr, 256, sw.out_proj, sw.attn_norm, R, xout[mode], 256, false, xin, 256, st,
                  tc ? (mode == 1 ? &hq_ : &hx1_) : nullptr, qkv_, T, mode);
        xin = xout[mode];
      }
      const float* attn = xb_;
      // DynamicConv (transformer.py:1116-1164)
      // with the tensor-core DynamicConv the parameters leave the GEMM as split-fp16 planes (TMA-store epilogue) in
      // the same buffer: hi plane, then lo plane
      Planes pp;
      pp.hi = reinterpret_cast<__half*>(params_);
      pp.lo = pp.hi + static_cast<size_t>(R) * 32768;
      const bool par_planes = dyn_mma_ && tc;
      linear_tc(sk + "dyn", attn, 256, hq_, tc, sw.dyn, R, params_, false, nullptr, st, 1, par_planes ? &pp : nullptr);
      if (fork) MCG_CUDA(cudaStreamWaitEvent(st, join_ev_[s], 0));
      if (dyn_mma_) {
        if (par_planes)
          dynconv_mma_kernel<true><<<R, 256, kDynMmaSmemBytes, st>>>(roih_.hi, roih_.lo, nullptr, pp.hi, pp.lo, sw.norm_in.g,
                                                                     sw.norm_in.b, sw.norm_out.g, sw.norm_out.b, dynf_,
                                                                     hf_.hi, hf_.lo);
        else
          dynconv_mma_kernel<false><<<R, 256, kDynMmaSmemBytes, st>>>(roih_.hi, roih_.lo, params_, nullptr, nullptr,
                                                                      sw.norm_in.g, sw.norm_in.b, sw.norm_out.g,
                                                                      sw.norm_out.b, dynf_, tc ? hf_.hi : nullptr,
                                                                      tc ? hf_.lo : nullptr);
      } else {
        dynconv_kernel<<<R, 256, kDynSmemBytes, st>>>(roi_, params_, sw.norm_in.g, sw.norm_in.b, sw.norm_out.g,
                                                      sw.norm_out.b, dynf_, tc ? hf_.hi : nullptr, tc ? hf_.lo : nullptr);
      }
      MCG_CUDA(cudaGetLastError());
      count(dyn_mma_ ? "dynconv_mma_kernel" : "dynconv_kernel");
      {
        // 12544 -> 256 over only 3*frames rows: split K so that >100 tiles exist; the partial sums are
        // reduced (and the bias added) inside the fc_norm LayerNorm kernel
        const int ks = tc ? kFcSplit : 1;
        linear_tc(sk + "fc", dynf_, 12544, hf_, tc, sw.fc, R, fcp_, false, nullptr, st, ks);
        // fc_norm + ReLU (transformer.py:1160-1162), then obj = LN(attn + iic) (gaze_stqi_head.py:175-176): one kernel
        ln(fcp_, 256, nullptr, 0, sw.fc_norm, xa_, 256, R, true, st, ks, static_cast<long long>(R) * 256,
           ks > 1 ? sw.fc.bias : nullptr, tc ? &hq_ : nullptr, attn, 256, &sw.iic_norm);
      }
      // FFN with identity (gaze_stqi_head.py:179): the hidden activations stay split-fp16 planes
      linear_tc(sk + "ffn1", xa_, 256, hq_, tc, sw.ffn1, R, ffn_h_, true, nullptr, st, 1, tc ? &hh_ : nullptr);
      {
        // 256 x 2048 over 3 frames rows is 24 tiles with 32 k-blocks each: split K, the ffn_norm LayerNorm reduces the
        // partial sums and adds bias + identity (gaze_stqi_head.py:179); it also emits the next stage's in_proj operand
        const int ks = tc ? kFfnSplit : 1;
        linear_tc(sk + "ffn2", ffn_h_, 2048, hh_, tc, sw.ffn2, R, ks > 1 ? ffn_p_ : xc_, false, ks > 1 ? nullptr : xa_, st, ks);
        if (ks > 1)
          ln(ffn_p_, 256, xa_, 256, sw.ffn_norm, obj_out, 256, R, false, st, ks, static_cast<long long>(R) * 256, sw.ffn2.bias,
             &hobj_);
        else
          ln(xc_, 256, nullptr, 0, sw.ffn_norm, obj_out, 256, R, false, st);
      }
      // cls / reg towers + per-clue heads (gaze_stqi_head.py:185-201); the cls tower and the first reg layer
      // read the same input and run as one grouped launch, as do the three per-clue heads of each kind
      // The classification branch only feeds the detection scores, and only the last stage's reach the output
      // (multiclue_gaze_roi_head.py:351-366): stages 0-2 skip it unless their intermediates are being recorded.
      const bool need_cls = s == 3 || (keep_stage_interm_ && !graph_mode_);
      {
        // one launch: chain 0 = reg tower (3 x Linear + LN + ReLU) -> per-clue fc_reg -> delta2bbox, chain 1 = cls
        // tower (1 layer) -> per-clue fc_cls
        ChainGroups cg = {};
        ChainArgs& r = cg.g[0];
        r.x = obj_out;
        r.ldx = 256;
        r.n_layers = 3;
        for (int j = 0; j < 3; ++j) {
          MCG_CHECK(sw.reg_fc[j].w_t != nullptr && sw.reg_fc[j].bias == nullptr, "reg tower layout");
          r.wt[j] = sw.reg_fc[j].w_t;
          r.gamma[j] = sw.reg_ln[j].g;
          r.beta[j] = sw.reg_ln[j].b;
          r.fw[j] = sw.fc_reg[j].w_f32;
          r.fb[j] = sw.fc_reg[j].bias;
        }
        r.n_classes = 3;
        r.nout = 4;
        r.y = delta_;
        r.ldy = 4;
        r.boxes_in = boxes_in;
        r.boxes_out = boxes_out;
        ChainArgs& c = cg.g[1];
        c.x = obj_out;
        c.ldx = 256;
        c.n_layers = 1;
        MCG_CHECK(sw.cls_fc.w_t != nullptr && sw.cls_fc.bias == nullptr, "cls tower layout");
        c.wt[0] = sw.cls_fc.w_t;
        c.gamma[0] = sw.cls_ln.g;
        c.beta[0] = sw.cls_ln.b;
        for (int j = 0; j < 3; ++j) {
          c.fw[j] = sw.fc_cls[j].w_f32;
          c.fb[j] = sw.fc_cls[j].bias;
        }
        c.n_classes = 3;
        c.nout = 1;
        c.y = cls_logit_;
        c.ldy = 1;
        dim3 grid(static_cast<unsigned>((R + kSlRows - 1) / kSlRows), need_cls ? 2 : 1);
        mlp_chain_kernel<<<grid, 1024, kChainSmemBytes, st>>>(cg, R);
        MCG_CUDA(cudaGetLastError());
        count("mlp_chain_kernel");
      }
      if (keep_stage_interm_ && !graph_mode_) {
        // head buffers are reused by every stage: snapshot them for per-op parity tests
        const std::string nm = "stage" + std::to_string(s);
        if (dyn_mma_) {  // the RoI features live as planes: export them as fp32
          planes_to_f32_kernel<<<1024, 256, 0, st>>>(roih_.hi, roih_.lo, nullptr, static_cast<long long>(R) * 12544, roi_);
          MCG_CUDA(cudaGetLastError());
        }
        reg_f32(nm + ".roi_feat", snapshot(roi_, static_cast<size_t>(R) * 12544, st), R, 49, 256);
        reg_f32(nm + ".attn", snapshot(attn, static_cast<size_t>(R) * 256, st), NB, 3, 256);
        reg_f32(nm + ".obj", snapshot(obj_out, static_cast<size_t>(R) * 256, st), NB, 3, 256);
        reg_f32(nm + ".boxes", snapshot(boxes_out, static_cast<size_t>(R) * 4, st), NB, 3, 4);
        reg_f32(nm + ".delta", snapshot(delta_, static_cast<size_t>(R) * 4, st), NB, 3, 4);
        reg_f32(nm + ".cls", snapshot(cls_logit_, static_cast<size_t>(R), st), NB, 3);
      }
      cur ^= 1;
    }
    reg_f32("obj", obj_[cur], NB, 3, 256);
    reg_f32("boxes", boxes_[cur], NB, 3, 4);
    reg_f32("cls", cls_logit_, NB, 3);
    // ---- gaze head on the last stage's object features (gaze_head.py:138-202)
    // the 3 clues x {gaze, confidence} branches are six independent tower chains (2 x Linear + LN + ReLU -> fc): one launch
    const float* obj = obj_[cur];
    {
      ChainGroups cg = {};
      for (int c = 0; c < 3; ++c) {
        for (int branch = 0; branch < 2; ++branch) {
          ChainArgs& a = cg.g[c * 2 + branch];
          const GemmW* tw = branch == 0 ? gaze_.tower[c] : gaze_.ctower[c];
          const LnW* tl = branch == 0 ? gaze_.tower_ln[c] : gaze_.ctower_ln[c];
          const GemmW& fc = branch == 0 ? gaze_.fc[c] : gaze_.fc_conf[c];
          a.x = obj + c * 256;   // rows of clue c: [NB] rows with stride 768
          a.ldx = 768;
          a.n_layers = 2;
          for (int j = 0; j < 2; ++j) {
            MCG_CHECK(tw[j].w_t != nullptr && tw[j].bias == nullptr, "gaze tower layout");
            a.wt[j] = tw[j].w_t;
            a.gamma[j] = tl[j].g;
            a.beta[j] = tl[j].b;
          }
          a.n_classes = 1;
          a.fw[0] = fc.w_f32;
          a.fb[0] = fc.bias;
          a.nout = 3;
          a.y = (branch == 0 ? gvec_ : conf_) + static_cast<size_t>(c) * NB * 3;
          a.ldy = 3;
        }
      }
      dim3 grid(static_cast<unsigned>((NB + kSlRows - 1) / kSlRows), 6);
      mlp_chain_kernel<<<grid, 1024, kChainSmemBytes, st>>>(cg, NB);
      MCG_CUDA(cudaGetLastError());
      count("mlp_chain_kernel");
    }
    finalize_kernel<<<(NB + 127) / 128, 128, 0, st>>>(gvec_, conf_, gaze_.wg, gaze_.bg, cls_logit_, boxes_[cur],
                                                      has_scale_ ? scale : nullptr, NB, out_gaze, out_boxes, out_scores);
    MCG_CUDA(cudaGetLastError());
    count("finalize_kernel");
  }

  // -------------------------------------------------------------------------- state
  int device_ = 0;
  int precision_ = 0;
  int num_sms_ = 148;
  bool head_tc_ = true;
  bool fused_stem_ = true;
  bool fuse_ds_ = true;  // conv3 + downsample branch of a layer's first bottleneck as one K-concatenated GEMM
  // consecutive tcgen05 launches walk their tile lists in opposite directions (env MCG_TUNE_NO_ALTERNATE=1: all forward)
  bool alternate_ = std::getenv("MCG_TUNE_NO_ALTERNATE") == nullptr;
  int tc_launch_index_ = 0;
  bool fuse_bneck_ = true;  // conv2 -> conv3 + identity of the other bottlenecks of layer1 / layer2 as one kernel (fp16c8)
  // Layers (bit l = layer l+1) whose bottlenecks run as TWO chains, one per half of the frames, on two streams: with
  // 172 (layer3) / 86-172 (layer4) tiles for 74 SM pairs a single persistent launch leaves a quarter of the machine idle in
  // its last wave (ncu: sm__cycles_active min / max = 0.70-0.75 over the SMs); the other half's launches fill it.  Frames
  // are independent through the trunk and a tile's k order does not depend on the tiling, so the results are bit-identical.
  // Option "split_layers" / env MCG_TUNE_SPLIT_LAYERS; batches below split_min_frames_ keep one chain.
  static int default_split_layers() {
    const char* e = std::getenv("MCG_TUNE_SPLIT_LAYERS");
    return e ? (std::atoi(e) & 0xf) : kDefaultSplitLayers;
  }
  static int default_split_min_frames() {
    const char* e = std::getenv("MCG_TUNE_SPLIT_MIN_NB");
    return e ? std::atoi(e) : 64;
  }
  int split_layers_ = default_split_layers();
  int split_min_frames_ = default_split_min_frames();
  int reverse_override_ = -1;  // >= 0: tile order of the next plan (the two chains alternate on their own)
  StemFusedPlan stem_plan_;
  bool stem_plan_valid_ = false;
  bool keep_stage_interm_ = false;
  bool time_kernels_ = false;
  std::vector<cudaEvent_t> prof_ev_;
  size_t prof_used_ = 0;
  std::vector<std::string> prof_names_;
  cudaStream_t cur_stream_ = nullptr;
  std::vector<cudaEvent_t> ev_pool_;
  size_t ev_used_ = 0;
  int umma_launches_ = 0;
  double umma_flops_ = 0.0;
  std::vector<std::unique_ptr<DeviceBlock>> dbg_;
  std::vector<float> meta_host_;
  std::unordered_map<std::string, HostT> host_;
  std::vector<std::unique_ptr<DeviceBlock>> keep_;
  std::vector<RangeRow> weight_ranges_;
  ConvW stem_;
  std::vector<BlockW> blocks_[4];
  ConvW lateral_[4], fpnconv_[4];
  const float* init_boxes_ = nullptr;
  const float* init_feats_ = nullptr;
  StageW stage_[4];
  GazeW gaze_;

  Arena arena_;
  int ws_NB_ = 0, ws_T_ = 0, ws_H_ = 0, ws_W_ = 0;
  struct BlkAct {
    Act t1, t2, ds, out;
  };
  Planes stemA_;
  Act stem_out_, pool_out_;
  std::vector<BlkAct> blk_act_[4];
  Act lat_[4], fpn_[4];
  float *boxes_[2] = {nullptr, nullptr}, *obj_[2] = {nullptr, nullptr};
  float *qkv_ = nullptr, *att_ = nullptr, *xa_ = nullptr, *xb_ = nullptr, *xc_ = nullptr, *params_ = nullptr,
        *roi_ = nullptr, *dynf_ = nullptr, *fc_ = nullptr, *fcp_ = nullptr, *ffn_h_ = nullptr, *t256a_ = nullptr, *t256b_ = nullptr,
        *cls_logit_ = nullptr, *delta_ = nullptr, *gz_a_ = nullptr, *gz_b_ = nullptr, *gvec_ = nullptr,
        *conf_ = nullptr, *d_meta_ = nullptr;
  Planes hq_, hh_, hf_, roih_, hobj_, hx1_;
  float* ffn_p_ = nullptr;  // split-K partial sums of the second FFN Linear
  bool dyn_mma_ = false;
  static constexpr int kMetaSlots = 4;
  float* pin_meta_ = nullptr;                 // kMetaSlots x [NB * 6] pinned staging of the per-call metadata
  size_t pin_meta_bytes_ = 0;
  cudaEvent_t meta_ev_[kMetaSlots] = {};
  int meta_slot_ = 0;
  bool has_scale_ = false;

  std::map<std::string, UmmaPlan> plans_;
  std::map<std::string, BneckPlan> bneck_plans_;
  std::map<std::string, Interm> interm_;
  int launches_ = 0;

  bool graph_mode_ = false;
  struct GraphKey {
    const float* img;
    float *gaze, *boxes, *scores;
    bool has_scale;
    bool operator<(const GraphKey& o) const {
      return std::tie(img, gaze, boxes, scores, has_scale) < std::tie(o.img, o.gaze, o.boxes, o.scores, o.has_scale);
    }
  };
  std::map<GraphKey, cudaGraphExec_t> graphs_;
  struct ShapeCtx {  // launch plans / graphs of a (NB, T, H, W) seen before (valid while the arena is not re-allocated)
    std::map<std::string, UmmaPlan> plans;
    std::map<std::string, BneckPlan> bneck_plans;
    std::map<GraphKey, cudaGraphExec_t> graphs;
    StemFusedPlan stem_plan;
    bool stem_valid = false;
  };
  std::map<std::tuple<int, int, int, int>, ShapeCtx> ctx_;

  cudaStream_t own_stream_ = nullptr;
  cudaStream_t side_stream_ = nullptr;  // forked branch of the head (RoIAlign beside the attention block)
  cudaEvent_t fork_ev_[4] = {nullptr, nullptr, nullptr, nullptr}, join_ev_[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t split_fork_ev_ = nullptr, split_join_ev_ = nullptr;  // two-chain region of the trunk (split_layers_)
  cudaStream_t cap_stream_ = nullptr;
  cudaStream_t copy_stream_ = nullptr;
  struct HostSlot {
    std::unique_ptr<DeviceBlock> d_in, d_out;
    size_t in_bytes = 0, out_bytes = 0;
    void* pin_out = nullptr;
    cudaEvent_t h2d_done = nullptr, done = nullptr;
    bool busy = false;
    int NB = 0;
  };
  HostSlot slot_[2];
  int next_slot_ = 0;
};

}  // namespace mcg

namespace mcg {
// preprocess.cu
void preprocess_launch(const mcg_frame* frames, int n, const float* mean, const float* std_, int to_rgb, float* out,
                       int Hp, int Wp, cudaStream_t st, int* launches);
// png_decode.cu
int png_parse(const uint8_t* f, int64_t n, int check_crc, mcg_png_info* info, uint8_t* z, int64_t zcap, const char** why);
void png_decode_launch(const mcg_png_job* jobs, int n, int32_t* status, cudaStream_t st, int* launches);
void png_file_sizes(const char* const* paths, int n, int64_t* sizes);
int png_stage_files(const char* const* paths, int n, int check_crc, int threads, const int64_t* slot_off,
                    const int64_t* slot_cap, uint8_t* block, mcg_png_info* infos, int32_t* results);
// metric.cu
void gaze_error_launch(const float* pred, const float* gt, const int32_t* video_start, int n_videos, int variant,
                       double* out, cudaStream_t st);
void merge_clips_launch(const float* rows, const int32_t* clip_start, const int32_t* frame_start, int n_videos, int clip_len,
                        int stride, float* det, float* gaze, cudaStream_t st);
}  // namespace mcg

// ========================================================================================
// C ABI
// ========================================================================================
struct mcg_engine {
  std::unique_ptr<mcg::Engine> impl;
};

template <typename F>
static int guarded(F&& f) {
  try {
    return f();
  } catch (const mcg::CudaError& e) {
    mcg::g_last_error = e.what();
    const std::string m = e.what();
    if (m.find("missing checkpoint key") != std::string::npos) return MCG_ERR_MISSING_WEIGHT;
    if (m.find("check failed") != std::string::npos) return MCG_ERR_INVALID;
    return MCG_ERR_CUDA;
  } catch (const std::exception& e) {
    mcg::g_last_error = e.what();
    return MCG_ERR_INVALID;
  } catch (...) {
    mcg::g_last_error = "unknown error";
    return MCG_ERR_INVALID;
  }
}

extern "C" {

const char* mcg_last_error(void) { return mcg::g_last_error.c_str(); }
const char* mcg_version(void) { return "mcgaze_b200 0.1 (sm_100a)"; }

int mcg_create(mcg_handle* out, int device, const mcg_tensor* weights, int n_weights, int precision) {
  return guarded([&]() -> int {
    if (!out || !weights || n_weights <= 0 || precision < 0 || precision > MCG_PRECISION_FP16C8) {
      mcg::g_last_error = "mcg_create: invalid argument";
      return MCG_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      mcg::g_last_error = "mcg_create: no CUDA device visible (this library has no CPU fallback)";
      return MCG_ERR_CUDA;
    }
    std::unique_ptr<mcg_engine> h(new mcg_engine);
    h->impl.reset(new mcg::Engine(device, weights, n_weights, precision));
    *out = h.release();
    return MCG_OK;
  });
}

int mcg_destroy(mcg_handle h) {
  return guarded([&]() -> int {
    delete h;
    return MCG_OK;
  });
}

int mcg_forward(mcg_handle h, const float* img, int B, int T, int H, int W, const float* img_hw,
                const float* scale_factor, float* out_gaze, float* out_boxes, float* out_scores, void* stream) {
  return guarded([&]() -> int {
    if (!h || !img || !out_gaze || !out_boxes || !out_scores) {
      mcg::g_last_error = "mcg_forward: null argument";
      return MCG_ERR_INVALID;
    }
    h->impl->forward(img, B, T, H, W, img_hw, scale_factor, out_gaze, out_boxes, out_scores,
                     static_cast<cudaStream_t>(stream));
    return MCG_OK;
  });
}

int mcg_forward_host(mcg_handle h, const float* img_host, int B, int T, int H, int W, const float* img_hw,
                     const float* scale_factor, float* out_gaze_host, float* out_boxes_host, float* out_scores_host) {
  return guarded([&]() -> int {
    if (!h || !img_host || !out_gaze_host || !out_boxes_host || !out_scores_host) {
      mcg::g_last_error = "mcg_forward_host: null argument";
      return MCG_ERR_INVALID;
    }
    h->impl->forward_host(img_host, B, T, H, W, img_hw, scale_factor, out_gaze_host, out_boxes_host, out_scores_host);
    return MCG_OK;
  });
}

int mcg_submit_host(mcg_handle h, const float* img_host, int B, int T, int H, int W, const float* img_hw,
                    const float* scale_factor, int* ticket) {
  return guarded([&]() -> int {
    if (!h || !img_host || !ticket) {
      mcg::g_last_error = "mcg_submit_host: null argument";
      return MCG_ERR_INVALID;
    }
    *ticket = h->impl->submit_host(img_host, B, T, H, W, img_hw, scale_factor);
    return MCG_OK;
  });
}

int mcg_wait_host(mcg_handle h, int ticket, float* out_gaze_host, float* out_boxes_host, float* out_scores_host) {
  return guarded([&]() -> int {
    if (!h || !out_gaze_host || !out_boxes_host || !out_scores_host) {
      mcg::g_last_error = "mcg_wait_host: null argument";
      return MCG_ERR_INVALID;
    }
    h->impl->wait_host(ticket, out_gaze_host, out_boxes_host, out_scores_host);
    return MCG_OK;
  });
}

int mcg_preprocess(const mcg_frame* frames, int n, const float* mean, const float* std, int to_rgb, float* out,
                   int Hp, int Wp, void* stream) {
  return guarded([&]() -> int {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      mcg::g_last_error = "mcg_preprocess: no CUDA device visible (this library has no CPU fallback)";
      return MCG_ERR_CUDA;
    }
    mcg::preprocess_launch(frames, n, mean, std, to_rgb, out, Hp, Wp, static_cast<cudaStream_t>(stream), nullptr);
    return MCG_OK;
  });
}

int mcg_png_parse(const uint8_t* file, int64_t nbytes, int check_crc, mcg_png_info* info, uint8_t* zdata, int64_t zcap) {
  return guarded([&]() -> int {
    const char* why = "";
    const int rc = mcg::png_parse(file, nbytes, check_crc, info, zdata, zcap, &why);
    if (rc != MCG_OK) mcg::g_last_error = std::string("mcg_png_parse: ") + why;
    return rc;
  });
}

int mcg_png_file_sizes(const char* const* paths, int n, int64_t* sizes) {
  return guarded([&]() -> int {
    if (!paths || !sizes || n <= 0) {
      mcg::g_last_error = "mcg_png_file_sizes: null argument";
      return MCG_ERR_INVALID;
    }
    mcg::png_file_sizes(paths, n, sizes);
    return MCG_OK;
  });
}

int mcg_png_stage_files(const char* const* paths, int n, int check_crc, int threads, const int64_t* slot_off,
                        const int64_t* slot_cap, uint8_t* block, mcg_png_info* infos, int32_t* results) {
  return guarded([&]() -> int {
    if (!paths || !slot_off || !slot_cap || !block || !infos || !results || n <= 0) {
      mcg::g_last_error = "mcg_png_stage_files: null argument";
      return MCG_ERR_INVALID;
    }
    const int bad = mcg::png_stage_files(paths, n, check_crc, threads, slot_off, slot_cap, block, infos, results);
    if (bad) {
      for (int i = 0; i < n; ++i)
        if (results[i]) {
          static const char* what[] = {"", "cannot read", "not a PNG file or malformed", "a PNG the device decoder does not take"};
          mcg::g_last_error = std::string("mcg_png_stage_files: ") + (paths[i] ? paths[i] : "(null)") + ": " + what[results[i] & 3] +
                              " (" + std::to_string(bad) + " of " + std::to_string(n) + " files)";
          break;
        }
      return MCG_ERR_INVALID;
    }
    return MCG_OK;
  });
}

int mcg_png_decode(const mcg_png_job* jobs, int n, int32_t* status, void* stream) {
  return guarded([&]() -> int {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      mcg::g_last_error = "mcg_png_decode: no CUDA device visible (this library has no CPU fallback)";
      return MCG_ERR_CUDA;
    }
    mcg::png_decode_launch(jobs, n, status, static_cast<cudaStream_t>(stream), nullptr);
    return MCG_OK;
  });
}

int mcg_merge_clips(const float* rows, const int32_t* clip_start, const int32_t* frame_start, int n_videos, int clip_len,
                    int stride, float* det, float* gaze, void* stream) {
  return guarded([&]() -> int {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      mcg::g_last_error = "mcg_merge_clips: no CUDA device visible (this library has no CPU fallback)";
      return MCG_ERR_CUDA;
    }
    mcg::merge_clips_launch(rows, clip_start, frame_start, n_videos, clip_len, stride, det, gaze,
                            static_cast<cudaStream_t>(stream));
    return MCG_OK;
  });
}

int mcg_gaze_error(const float* pred, const float* gt, const int32_t* video_start, int n_videos, int variant, double* out,
                   void* stream) {
  return guarded([&]() -> int {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      mcg::g_last_error = "mcg_gaze_error: no CUDA device visible (this library has no CPU fallback)";
      return MCG_ERR_CUDA;
    }
    mcg::gaze_error_launch(pred, gt, video_start, n_videos, variant, out, static_cast<cudaStream_t>(stream));
    return MCG_OK;
  });
}

int mcg_get_intermediate(mcg_handle h, const char* name, float* dst, int64_t capacity, int64_t shape_out[4]) {
  return guarded([&]() -> int {
    if (!h || !name || !dst || !shape_out) return MCG_ERR_INVALID;
    const int r = h->impl->get_intermediate(name, dst, capacity, shape_out);
    if (r != MCG_OK) mcg::g_last_error = std::string("mcg_get_intermediate: unknown name or capacity too small: ") + name;
    return r;
  });
}

int mcg_last_launch_count(mcg_handle h) { return h ? h->impl->last_launches() : -1; }

int mcg_last_umma_stats(mcg_handle h, double out[3]) {
  return guarded([&]() -> int {
    if (!h || !out) return MCG_ERR_INVALID;
    h->impl->umma_stats(out);
    return MCG_OK;
  });
}

int mcg_last_umma_times(mcg_handle h, double* out_ms, int capacity) {
  int n = -1;
  const int rc = guarded([&]() -> int {
    if (!h || !out_ms || capacity <= 0) return MCG_ERR_INVALID;
    n = h->impl->umma_times(out_ms, capacity);
    return MCG_OK;
  });
  return rc == MCG_OK ? n : rc;
}

int mcg_last_kernel_profile(mcg_handle h, char* buf, int capacity) {
  int n = -1;
  const int rc = guarded([&]() -> int {
    if (!h || (capacity > 0 && !buf)) return MCG_ERR_INVALID;
    const std::string s = h->impl->kernel_profile();
    n = static_cast<int>(s.size()) + 1;
    if (capacity >= n) std::memcpy(buf, s.c_str(), static_cast<size_t>(n));
    return MCG_OK;
  });
  return rc == MCG_OK ? n : rc;
}

int mcg_range_report(mcg_handle h, char* buf, int capacity) {
  int n = -1;
  const int rc = guarded([&]() -> int {
    if (!h || (capacity > 0 && !buf)) return MCG_ERR_INVALID;
    const std::string s = h->impl->range_report();
    n = static_cast<int>(s.size()) + 1;
    if (capacity >= n) std::memcpy(buf, s.c_str(), static_cast<size_t>(n));
    return MCG_OK;
  });
  return rc == MCG_OK ? n : rc;
}

int mcg_set_graph_mode(mcg_handle h, int on) {
  return guarded([&]() -> int {
    if (!h) return MCG_ERR_INVALID;
    h->impl->set_graph_mode(on);
    return MCG_OK;
  });
}

int mcg_set_option(mcg_handle h, const char* key, int value) {
  return guarded([&]() -> int {
    if (!h || !key) return MCG_ERR_INVALID;
    h->impl->set_option(key, value);
    return MCG_OK;
  });
}

int mcg_debug_conv(int engine, const float* x, int NB, int H, int W, int C, const float* w, int Cout, int R, int S,
                   int stride, int pad, const float* bias, const float* res, int res_mode, int relu, int force_im2col,
                   int force_block_n, int out_mode, float* out, void* stream) {
  using namespace mcg;
  return guarded([&]() -> int {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
    const long long M = static_cast<long long>(NB) * P * Q;
    const int K = R * S * C;
    const size_t nx = static_cast<size_t>(NB) * H * W * C, nw = static_cast<size_t>(Cout) * K;
    const size_t ny = static_cast<size_t>(M) * Cout;
    int RH = P, RW = Q;
    if (res_mode == RES_UP2X) {
      RH = P / 2;
      RW = Q / 2;
    }
    const size_t nr = res ? static_cast<size_t>(NB) * RH * RW * Cout : 0;
    // fp16c8: out_mode bit 2 = return the emitted hi8 output plane (as fp32) instead of hi + lo8
    const bool c8 = engine == MCG_PRECISION_FP16C8;
    const bool c8_hi8_out = c8 && (out_mode & 4);
    out_mode &= 1;
    DeviceBlock bx(nx * 4), bw(nw * 4), br(nr * 4 + 16);
    DeviceBlock bx8(c8 ? nx * 2 : 16), bw8(c8 ? nw * 2 : 16), br8(c8 ? nr + 16 : 16);
    Planes px{reinterpret_cast<__half*>(bx.p), reinterpret_cast<__half*>(bx.p) + nx, nullptr};
    Planes pw{reinterpret_cast<__half*>(bw.p), reinterpret_cast<__half*>(bw.p) + nw, nullptr};
    Planes pr{reinterpret_cast<__half*>(br.p), reinterpret_cast<__half*>(br.p) + nr, nullptr};
    if (c8) {
      px.lo8 = reinterpret_cast<uint8_t*>(bx8.p);
      px.hi8 = px.lo8 + nx;
      pr.lo8 = reinterpret_cast<uint8_t*>(br8.p);
    }
    split_planes_kernel<<<1024, 256, 0, st>>>(x, C, static_cast<long long>(NB) * H * W, C, px.hi, px.lo, px.lo8, px.hi8);
    split_planes_kernel<<<1024, 256, 0, st>>>(w, K, Cout, K, pw.hi, pw.lo);
    if (res)
      split_planes_kernel<<<1024, 256, 0, st>>>(res, Cout, static_cast<long long>(NB) * RH * RW, Cout, pr.hi, pr.lo, pr.lo8);
    MCG_CUDA(cudaGetLastError());
    if (c8) {
      // same e4m3 weight packing as the engine (host side)
      std::vector<float> hw(nw);
      std::vector<__half> hh(nw);
      std::vector<uint8_t> hi8(nw), lo8(nw);
      MCG_CUDA(cudaMemcpyAsync(hw.data(), w, nw * 4, cudaMemcpyDeviceToHost, st));
      MCG_CUDA(cudaStreamSynchronize(st));
      for (size_t i = 0; i < nw; ++i) hh[i] = __float2half_rn(hw[i]);
      pack_w8_host(hw.data(), hh.data(), nw, hi8.data(), lo8.data());
      pw.hi8 = reinterpret_cast<uint8_t*>(bw8.p);
      pw.lo8 = pw.hi8 + nw;
      MCG_CUDA(cudaMemcpy(pw.hi8, hi8.data(), nw, cudaMemcpyHostToDevice));
      MCG_CUDA(cudaMemcpy(pw.lo8, lo8.data(), nw, cudaMemcpyHostToDevice));
    }
    AGeom g;
    const bool plain = R == 1 && S == 1 && stride == 1 && pad == 0 && !force_im2col;
    g.kind = plain ? 0 : 1;
    g.lda = C;
    g.NB = NB;
    g.H = H;
    g.W = W;
    g.C = C;
    g.R = R;
    g.S = S;
    g.stride = stride;
    g.pad = pad;
    g.P = P;
    g.Q = Q;
    // tensor-core engines write split-fp16 planes (the trunk's real output format: smem-staged TMA
    // stores); out_mode 1 forces the fp32 direct-store epilogue the head uses
    const bool planes_out = engine != MCG_PRECISION_SIMT && out_mode == 0;
    DeviceBlock bo(planes_out ? ny * 4 : 16);
    Planes po{reinterpret_cast<__half*>(bo.p), reinterpret_cast<__half*>(bo.p) + ny};
    Epilogue ep;
    ep.bias = bias;
    ep.relu = relu;
    if (planes_out) {
      ep.out_hi = po.hi;
      ep.out_lo = engine == MCG_PRECISION_FP16X3 ? po.lo : nullptr;
      if (c8) {
        ep.out_lo8 = reinterpret_cast<uint8_t*>(po.lo);
        ep.out_hi8 = c8_hi8_out ? ep.out_lo8 + ny : nullptr;
      }
    } else {
      ep.out_f32 = out;
    }
    ep.ldo = Cout;
    if (res) {
      ep.res_hi = pr.hi;
      ep.res_lo = (c8 || engine == MCG_PRECISION_FP16) ? nullptr : pr.lo;
      ep.res_lo8 = pr.lo8;
      ep.res_mode = res_mode;
      ep.ldr = Cout;
      ep.P = P;
      ep.Q = Q;
    }
    if (engine == MCG_PRECISION_SIMT) {
      SimtParams p;
      p.M = M;
      p.N = Cout;
      p.K = K;
      p.a = g;
      p.a_hi = px.hi;
      p.a_lo = px.lo;
      p.a_lo8 = px.lo8;
      p.w_hi = pw.hi;
      p.w_lo = pw.lo;
      p.ep = ep;
      launch_simt_gemm(p, st);
    } else {
      int dev = 0, sms = 148;
      MCG_CUDA(cudaGetDevice(&dev));
      MCG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      if (!umma_supported(M, Cout, K, g)) {
        g_last_error = "mcg_debug_conv: shape not supported by the tcgen05 kernel";
        return MCG_ERR_UNSUPPORTED;
      }
      const int terms = engine == MCG_PRECISION_FP16X3 ? 3 : c8 ? 2 : 1;
      // force_block_n >= 1000: the CTA-pair (cta_group::2) variant with block_n = force_block_n - 1000 (0 = automatic)
      const int pair = force_block_n >= 1000 ? 1 : 0;
      UmmaPlan pl = make_umma_plan(terms, px, g, pw, M, Cout, K, ep, sms, force_block_n % 1000, 1, 0, pair);
      launch_umma(pl, st);
      if (planes_out) {
        if (c8_hi8_out)
          e4m3_to_f32_kernel<<<1024, 256, 0, st>>>(ep.out_hi8, static_cast<long long>(ny), out);
        else
          planes_to_f32_kernel<<<1024, 256, 0, st>>>(po.hi, ep.out_lo, ep.out_lo8, static_cast<long long>(ny), out);
        MCG_CUDA(cudaGetLastError());
      }
    }
    MCG_CUDA(cudaStreamSynchronize(st));
    return MCG_OK;
  });
}

}  // extern "C"
