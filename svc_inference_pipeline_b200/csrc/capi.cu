// extern "C" surface of libbvg_b200.so (see include/bvg_b200.h) and the program runner.
#include <cstring>
#include <mutex>
#include <new>
#include <set>
#include <utility>
#include <vector>

#include "common.cuh"

namespace bvg {

// ---- error plumbing -------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static thread_local bool g_pdl = false;
bool& pdl_mode() { return g_pdl; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return BVG_ECUDA;
}

bool first_use_on_device(const void* kernel) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> seen;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  return seen.insert(std::make_pair(dev, kernel)).second;
}

// kernels (defined in the other translation units)
int amp_forward(const bvg_amp_desc* d, cudaStream_t st);
int conv_simt_forward(const bvg_conv_desc* d, cudaStream_t st);
int conv_umma_forward(const bvg_conv_desc* d, cudaStream_t st);
int post_forward(const bvg_post_desc* d, cudaStream_t st);
int pack_mel(const bvg_pack_desc* d, cudaStream_t st);
int tail_forward(const bvg_tail_desc* d, cudaStream_t st);
int stitch_forward(const bvg_stitch_desc* d, cudaStream_t st);
int logmel_forward(const bvg_logmel_desc* d, cudaStream_t st);
int rowop_forward(const bvg_rowop_desc* d, cudaStream_t st);
int diffembed_forward(const bvg_diffembed_desc* d, cudaStream_t st);
int sample_forward(const bvg_sample_desc* d, cudaStream_t st);
int convert(const bvg_tensor* src, const bvg_tensor* dst, size_t n, cudaStream_t st);
int conv_geometry(const bvg_conv_geom* g, bvg_conv_weights* w);
size_t conv_plane_elems(const bvg_conv_weights* w);
int pack_conv_weights(const bvg_conv_geom* g, const float* d_v, const float* d_g, const float* d_bias, bvg_conv_weights* w, float* d_bias_out,
                      float* d_scale_scratch, cudaStream_t st);
int pack_post_weights(const float* d_v, const float* d_g, int cin, int ksize, float* d_w_out, float* d_scale_scratch, cudaStream_t st);

struct PairLaunch;
bool conv_pair_eligible(const bvg_conv_desc* d);
int conv_pair_prepare(const bvg_conv_desc* d, PairLaunch* out);
int conv_pair_launch(const PairLaunch* l, cudaStream_t st);
int conv_pair_forward(const bvg_conv_desc* d, cudaStream_t st);
size_t pair_launch_size();
struct UmmaLaunch;
int conv_umma_prepare(const bvg_conv_desc* d, UmmaLaunch* out);
int conv_umma_launch(const UmmaLaunch* l, cudaStream_t st);
size_t umma_launch_size();

static int conv_forward(const bvg_conv_desc* d, cudaStream_t st) {
  BVG_REQUIRE(d && d->w, "conv: null descriptor");
  BVG_REQUIRE(!d->pre_amp || d->w->backend == BVG_UMMA, "conv: a fused Activation1d (pre_amp) needs the UMMA backend");
  if (d->w->backend == BVG_SIMT) return conv_simt_forward(d, st);
  if (d->w->backend == BVG_UMMA) return conv_pair_eligible(d) ? conv_pair_forward(d, st) : conv_umma_forward(d, st);
  set_error("conv: unknown backend %d", d->w->backend);
  return BVG_EINVAL;
}

}  // namespace bvg

// A program owns deep copies of its op descriptors (and of the weight descriptors they point
// to), plus the pre-encoded TMA launch records of its tensor-core ops.
struct bvg_program {
  std::vector<bvg_op> ops;
  std::vector<bvg_conv_weights*> owned_weights;
  std::vector<bvg_tuning*> owned_tunings;   // deep copies of the descriptors' bvg_tuning (the caller's may go away)
  std::vector<void*> umma;  // UmmaLaunch* per op (nullptr for the others)
  std::vector<void*> pair;  // PairLaunch* per op: convolutions that run on the CTA-pair kernel
  int launches = 0;
  bool pdl = false;  // bvg_program_set_pdl: chain the launches with programmatic dependent launch (common.cuh)
  ~bvg_program() {
    for (auto* w : owned_weights) delete w;
    for (auto* t : owned_tunings) delete t;
    for (auto* u : umma) ::operator delete(u);
    for (auto* u : pair) ::operator delete(u);
  }
};

extern "C" {

int bvg_abi_version(void) { return BVG_ABI_VERSION; }
const char* bvg_last_error(void) { return bvg::g_err; }
size_t bvg_sizeof_op(void) { return sizeof(bvg_op); }
size_t bvg_sizeof_conv_weights(void) { return sizeof(bvg_conv_weights); }

int bvg_device_check(int device) {
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e != cudaSuccess) return bvg::cuda_fail(e, "cudaDeviceGetAttribute");
  if (major != 10) {
    bvg::set_error("device %d is sm_%d%d; libbvg_b200 is built for sm_100a only (no fallback)", device, major, minor);
    return BVG_EARCH;
  }
  return BVG_OK;
}

void bvg_tuning_defaults(bvg_tuning* t) {
  if (t) *t = bvg::default_tuning();
}

int bvg_amp_fwd(const bvg_amp_desc* d, void* stream) { return bvg::amp_forward(d, (cudaStream_t)stream); }
int bvg_conv_fwd(const bvg_conv_desc* d, void* stream) { return bvg::conv_forward(d, (cudaStream_t)stream); }
int bvg_post_fwd(const bvg_post_desc* d, void* stream) { return bvg::post_forward(d, (cudaStream_t)stream); }
int bvg_pack_mel(const bvg_pack_desc* d, void* stream) { return bvg::pack_mel(d, (cudaStream_t)stream); }
int bvg_tail_fwd(const bvg_tail_desc* d, void* stream) { return bvg::tail_forward(d, (cudaStream_t)stream); }
int bvg_rowop_fwd(const bvg_rowop_desc* d, void* stream) { return bvg::rowop_forward(d, (cudaStream_t)stream); }
int bvg_diffembed_fwd(const bvg_diffembed_desc* d, void* stream) { return bvg::diffembed_forward(d, (cudaStream_t)stream); }
int bvg_sample_fwd(const bvg_sample_desc* d, void* stream) { return bvg::sample_forward(d, (cudaStream_t)stream); }
int bvg_logmel_fwd(const bvg_logmel_desc* d, void* stream) { return bvg::logmel_forward(d, (cudaStream_t)stream); }
int bvg_stitch_fwd(const bvg_stitch_desc* d, void* stream) { return bvg::stitch_forward(d, (cudaStream_t)stream); }
int bvg_convert(const bvg_tensor* src, const bvg_tensor* dst, size_t n, void* stream) {
  return bvg::convert(src, dst, n, (cudaStream_t)stream);
}

int bvg_conv_geometry(const bvg_conv_geom* g, bvg_conv_weights* w) { return bvg::conv_geometry(g, w); }

int bvg_conv_pack_bytes(const bvg_conv_geom* g, size_t* weight_plane_bytes, size_t* bias_bytes) {
  bvg_conv_weights w;
  memset(&w, 0, sizeof(w));
  int rc = bvg::conv_geometry(g, &w);
  if (rc != BVG_OK) return rc;
  const size_t elems = bvg::conv_plane_elems(&w);
  if (weight_plane_bytes) *weight_plane_bytes = elems * (w.backend == BVG_SIMT ? sizeof(float) : 2);
  if (bias_bytes) *bias_bytes = (size_t)w.n_tiles * w.n_tile * sizeof(float);
  return BVG_OK;
}

int bvg_pack_conv_weights(const bvg_conv_geom* g, const float* d_v, const float* d_g, const float* d_bias, bvg_conv_weights* w,
                          float* d_bias_out, float* d_scratch, void* stream) {
  return bvg::pack_conv_weights(g, d_v, d_g, d_bias, w, d_bias_out, d_scratch, (cudaStream_t)stream);
}

int bvg_pack_post_weights(const float* d_v, const float* d_g, int32_t cin, int32_t ksize, float* d_w_out, float* d_scratch, void* stream) {
  return bvg::pack_post_weights(d_v, d_g, cin, ksize, d_w_out, d_scratch, (cudaStream_t)stream);
}

// ---- programs ---------------------------------------------------------------------------------
int bvg_program_create(const bvg_op* ops, int32_t n_ops, bvg_program** out) {
  if (!ops || n_ops < 0 || !out) {
    bvg::set_error("program_create: bad argument");
    return BVG_EINVAL;
  }
  bvg_program* p = new (std::nothrow) bvg_program();
  if (!p) {
    bvg::set_error("program_create: out of host memory");
    return BVG_ENOMEM;
  }
  p->ops.assign(ops, ops + n_ops);
  p->umma.assign(n_ops, nullptr);
  p->pair.assign(n_ops, nullptr);
  for (int i = 0; i < n_ops; ++i) {
    bvg_op& op = p->ops[i];
    if (op.kind == BVG_OP_AMP && op.u.amp.tune) {
      p->owned_tunings.push_back(new bvg_tuning(*op.u.amp.tune));
      op.u.amp.tune = p->owned_tunings.back();
    }
    if (op.kind == BVG_OP_CONV && op.u.conv.tune) {
      p->owned_tunings.push_back(new bvg_tuning(*op.u.conv.tune));
      op.u.conv.tune = p->owned_tunings.back();
    }
    if (op.kind == BVG_OP_CONV) {
      if (!op.u.conv.w) {
        bvg::set_error("program_create: op %d has no weights", i);
        delete p;
        return BVG_EINVAL;
      }
      bvg_conv_weights* w = new bvg_conv_weights(*op.u.conv.w);
      p->owned_weights.push_back(w);
      op.u.conv.w = w;
      if (w->backend != BVG_UMMA && op.u.conv.pre_amp) {
        bvg::set_error("program_create: op %d fuses an Activation1d into a convolution that is not on the UMMA backend", i);
        delete p;
        return BVG_EINVAL;
      }
      if (w->backend == BVG_UMMA && bvg::conv_pair_eligible(&op.u.conv)) {
        void* rec = ::operator new(bvg::pair_launch_size());
        p->pair[i] = rec;
        int rc = bvg::conv_pair_prepare(&op.u.conv, reinterpret_cast<bvg::PairLaunch*>(rec));
        if (rc != BVG_OK) {
          delete p;
          return rc;
        }
      } else if (w->backend == BVG_UMMA) {
        void* rec = ::operator new(bvg::umma_launch_size());
        p->umma[i] = rec;
        int rc = bvg::conv_umma_prepare(&op.u.conv, reinterpret_cast<bvg::UmmaLaunch*>(rec));
        if (rc != BVG_OK) {
          delete p;
          return rc;
        }
      }
    } else if (op.kind != BVG_OP_PACK && op.kind != BVG_OP_AMP && op.kind != BVG_OP_POST && op.kind != BVG_OP_ROWOP && op.kind != BVG_OP_DIFFEMBED && op.kind != BVG_OP_SAMPLE) {
      bvg::set_error("program_create: op %d has unknown kind %d", i, op.kind);
      delete p;
      return BVG_EINVAL;
    }
  }
  p->launches = n_ops;
  *out = p;
  return BVG_OK;
}

static int run_one(bvg_program* p, size_t i, cudaStream_t st) {
  const bvg_op& op = p->ops[i];
  bvg::PdlScope scope(p->pdl);
  switch (op.kind) {
    case BVG_OP_PACK: return bvg::pack_mel(&op.u.pack, st);
    case BVG_OP_AMP: return bvg::amp_forward(&op.u.amp, st);
    case BVG_OP_CONV:
      if (p->pair[i]) return bvg::conv_pair_launch(reinterpret_cast<const bvg::PairLaunch*>(p->pair[i]), st);
      return p->umma[i] ? bvg::conv_umma_launch(reinterpret_cast<const bvg::UmmaLaunch*>(p->umma[i]), st) : bvg::conv_forward(&op.u.conv, st);
    case BVG_OP_POST: return bvg::post_forward(&op.u.post, st);
    case BVG_OP_ROWOP: return bvg::rowop_forward(&op.u.rowop, st);
    case BVG_OP_DIFFEMBED: return bvg::diffembed_forward(&op.u.diffembed, st);
    case BVG_OP_SAMPLE: return bvg::sample_forward(&op.u.sample, st);
    default: return BVG_EINVAL;
  }
}

int bvg_program_run(bvg_program* p, void* stream) {
  if (!p) {
    bvg::set_error("program_run: null program");
    return BVG_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (size_t i = 0; i < p->ops.size(); ++i) {
    int rc = run_one(p, i, st);
    if (rc != BVG_OK) return rc;
  }
  return BVG_OK;
}

int bvg_program_set_pdl(bvg_program* p, int on) {
  if (!p) {
    bvg::set_error("program_set_pdl: null program");
    return BVG_EINVAL;
  }
  p->pdl = on != 0;
  return BVG_OK;
}

int bvg_program_run_interleaved(bvg_program* const* progs, void* const* streams, int32_t n) {
  if (!progs || !streams || n <= 0) {
    bvg::set_error("program_run_interleaved: bad argument");
    return BVG_EINVAL;
  }
  size_t n_ops = 0;
  for (int k = 0; k < n; ++k) {
    if (!progs[k]) {
      bvg::set_error("program_run_interleaved: null program");
      return BVG_EINVAL;
    }
    if (progs[k]->ops.size() > n_ops) n_ops = progs[k]->ops.size();
  }
  for (size_t i = 0; i < n_ops; ++i)
    for (int k = 0; k < n; ++k) {
      if (i >= progs[k]->ops.size()) continue;
      int rc = run_one(progs[k], i, (cudaStream_t)streams[k]);
      if (rc != BVG_OK) return rc;
    }
  return BVG_OK;
}

int bvg_program_run_timed(bvg_program* p, void* stream, float* ms_by_kind, int32_t* n_by_kind, float* ms_per_op) {
  if (!p || !ms_by_kind || !n_by_kind) {
    bvg::set_error("program_run_timed: bad argument");
    return BVG_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = p->ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) BVG_CHECK_CUDA(cudaEventCreate(&e));
  int rc = BVG_OK;
  for (size_t i = 0; i < n && rc == BVG_OK; ++i) {
    cudaEventRecord(ev[i], st);
    rc = run_one(p, i, st);
  }
  cudaEventRecord(ev[n], st);
  cudaError_t e = cudaStreamSynchronize(st);
  for (int k = 0; k < BVG_N_OP_KINDS; ++k) {
    ms_by_kind[k] = 0.f;
    n_by_kind[k] = 0;
  }
  if (rc == BVG_OK && e == cudaSuccess) {
    for (size_t i = 0; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      const int k = p->ops[i].kind;
      if (ms_per_op) ms_per_op[i] = ms;
      if (k >= 0 && k < BVG_N_OP_KINDS) {
        ms_by_kind[k] += ms;
        n_by_kind[k] += 1;
      }
    }
  }
  for (auto& x : ev) cudaEventDestroy(x);
  if (rc != BVG_OK) return rc;
  if (e != cudaSuccess) return bvg::cuda_fail(e, "cudaStreamSynchronize (program_run_timed)");
  return BVG_OK;
}

int bvg_program_num_launches(const bvg_program* p) { return p ? p->launches : 0; }
void bvg_program_destroy(bvg_program* p) { delete p; }

}  // extern "C"
