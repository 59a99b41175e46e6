// fluxb200 — error plumbing, device queries and TMA tensor-map encoding.
#include <cudaTypedefs.h>

#include <string.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "internal.h"

namespace fb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error() { return g_last_error.c_str(); }
int fail(const std::string& msg) {
  g_last_error = msg;
  return -1;
}

// ------------------------------------------------------------------------------------------------
// launch accounting / profiling
// ------------------------------------------------------------------------------------------------
struct ProfRec {
  cudaEvent_t a, b;
  int kind;
  double flops, bytes;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
static unsigned long long g_launches[KK_COUNT] = {0};

void count_launch(int kind, int n) { __atomic_fetch_add(&g_launches[kind], static_cast<unsigned long long>(n), __ATOMIC_RELAXED); }
void count_launch_bulk(const unsigned long long* per_kind, int sign) {
  for (int k = 0; k < KK_COUNT; ++k) {
    if (sign >= 0) __atomic_fetch_add(&g_launches[k], per_kind[k], __ATOMIC_RELAXED);
    else           __atomic_fetch_sub(&g_launches[k], per_kind[k], __ATOMIC_RELAXED);
  }
}
void snapshot_launches(unsigned long long* per_kind) {
  for (int k = 0; k < KK_COUNT; ++k) per_kind[k] = __atomic_load_n(&g_launches[k], __ATOMIC_RELAXED);
}
bool profiling_enabled() { return __atomic_load_n(&g_prof_on, __ATOMIC_RELAXED); }

bool DeviceOnce::need() const {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  return ((__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> dev) & 1ull) == 0;
}
void DeviceOnce::done() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELEASE);
}

ProfScope::ProfScope(int k, double flops, double bytes, cudaStream_t st) : kind(k), stream(st), slot(-1) {
  if (!profiling_enabled()) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  if (!g_prof_pool.empty()) {
    r.a = g_prof_pool.back().first, r.b = g_prof_pool.back().second;
    g_prof_pool.pop_back();
  } else {
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  }
  r.kind = k, r.flops = flops, r.bytes = bytes;
  cudaEventRecord(r.a, st);
  g_prof_recs.push_back(r);
  slot = static_cast<int>(g_prof_recs.size()) - 1;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof_recs[slot].b, stream);
}

static int g_flag_qkrope = 1, g_flag_pair = -1, g_flag_dqm = -1, g_flag_attn = -1, g_flag_pdl = -1, g_flag_cl4 = -1,
           g_flag_graph = -1, g_flag_big = -1, g_flag_dqo = -1, g_flag_lnr = -1;
int get_flag(const char* name) {
  if (!strcmp(name, "qkrope_fusion")) return g_flag_qkrope;
  if (!strcmp(name, "dequant_mode")) {
    if (g_flag_dqm < 0) {
      const char* e = getenv("FLUXB200_DEQUANT_MODE");
      g_flag_dqm = e ? (atoi(e) & 3) : 0;
      if (g_flag_dqm > 2) g_flag_dqm = 0;
    }
    return g_flag_dqm;
  }
  if (!strcmp(name, "fused_dequant")) return get_flag("dequant_mode") == 2;
  if (!strcmp(name, "gemm_cl4")) {
    if (g_flag_cl4 < 0) {
      const char* e = getenv("FLUXB200_GEMM_CL4");
      g_flag_cl4 = (e && e[0] == '1') ? 1 : 0;  // off by default: measured slower than CTA pairs
    }
    return g_flag_cl4;
  }
  if (!strcmp(name, "step_graph")) {
    if (g_flag_graph < 0) {
      const char* e = getenv("FLUXB200_STEP_GRAPH");
      g_flag_graph = (e && e[0] == '0') ? 0 : 1;
    }
    return g_flag_graph;
  }
  if (!strcmp(name, "ln_reread")) {
    if (g_flag_lnr < 0) {
      const char* e = getenv("FLUXB200_LN_REREAD");
      g_flag_lnr = e ? (e[0] != '0') : 1;
    }
    return g_flag_lnr;
  }
  if (!strcmp(name, "dequant_overlap")) {
    if (g_flag_dqo < 0) {
      const char* e = getenv("FLUXB200_DEQUANT_OVERLAP");
      g_flag_dqo = (e && e[0] == '0') ? 0 : 1;
    }
    return g_flag_dqo;
  }
  if (!strcmp(name, "gemm_big")) {
    if (g_flag_big < 0) {
      const char* e = getenv("FLUXB200_GEMM_BIG");
      g_flag_big = e ? atoi(e) : 0;  // off by default: measured slower
    }
    return g_flag_big;
  }
  if (!strcmp(name, "pdl")) {
    if (g_flag_pdl < 0) {
      const char* e = getenv("FLUXB200_PDL");
      g_flag_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_flag_pdl;
  }
  if (!strcmp(name, "attn_variant")) {
    if (g_flag_attn < 0) {
      const char* e = getenv("FLUXB200_ATTN_VARIANT");
      g_flag_attn = e ? atoi(e) : 0;
    }
    return g_flag_attn;
  }
  if (!strcmp(name, "gemm_pair")) {
    if (g_flag_pair < 0) {
      const char* e = getenv("FLUXB200_GEMM_SINGLE_CTA");
      g_flag_pair = (e && e[0] == '1') ? 0 : 1;
    }
    return g_flag_pair;
  }
  return 0;
}

int num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

// cuTensorMapEncodeTiled is a driver API; resolve it through the runtime so the library has no
// link-time dependency on libcuda.so (it must load on a CPU-only build box).
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

static int encode_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  auto enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string s = "cuTensorMapEncodeTiled failed, CUresult=" + std::to_string(static_cast<int>(r)) + " rank=" +
                    std::to_string(rank) + " dims=";
    for (int i = 0; i < rank; ++i) s += std::to_string(dims[i]) + ",";
    s += " strides=";
    for (int i = 0; i < rank - 1; ++i) s += std::to_string(strides_bytes[i]) + ",";
    s += " box=";
    for (int i = 0; i < rank; ++i) s += std::to_string(box[i]) + ",";
    return fail(s);
  }
  return 0;
}

int encode_tmap_2d_raw(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                       uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  auto enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {outer_stride_bytes};
  cuuint32_t bdim[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled (raw) failed, CUresult=" + std::to_string(static_cast<int>(r)) + " inner=" +
                std::to_string(inner) + " outer=" + std::to_string(outer) + " stride=" +
                std::to_string(outer_stride_bytes) + " box=" + std::to_string(box_inner) + "x" + std::to_string(box_outer));
  return 0;
}

int encode_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                   uint32_t box_inner, uint32_t box_outer) {
  uint64_t dims[2] = {inner, outer};
  uint64_t str[1] = {outer_stride_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return encode_nd(out, base, 2, dims, str, box);
}
int encode_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  uint64_t dims[3] = {d0, d1, d2};
  uint64_t str[2] = {stride1_bytes, stride2_bytes};
  uint32_t box[3] = {b0, b1, b2};
  return encode_nd(out, base, 3, dims, str, box);
}
int encode_tmap_4d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                   uint64_t stride1_bytes, uint64_t stride2_bytes, uint64_t stride3_bytes, uint32_t b0, uint32_t b1,
                   uint32_t b2, uint32_t b3) {
  uint64_t dims[4] = {d0, d1, d2, d3};
  uint64_t str[3] = {stride1_bytes, stride2_bytes, stride3_bytes};
  uint32_t box[4] = {b0, b1, b2, b3};
  return encode_nd(out, base, 4, dims, str, box);
}

}  // namespace fb

// ---- C ABI: profiling / accounting ----
extern "C" {
int fluxb200_set_flag(const char* name, int value) {
  if (!name) return fb::fail("set_flag: null name");
  if (!strcmp(name, "qkrope_fusion")) { fb::g_flag_qkrope = value ? 1 : 0; return 0; }
  if (!strcmp(name, "gemm_pair")) { fb::g_flag_pair = value ? 1 : 0; return 0; }
  if (!strcmp(name, "dequant_mode")) {
    if (value < 0 || value > 2) return fb::fail("set_flag: dequant_mode must be 0 (per-image cache), 1 (staged) or 2 (fused)");
    fb::g_flag_dqm = value;
    return 0;
  }
  if (!strcmp(name, "fused_dequant")) { fb::g_flag_dqm = value ? 2 : 0; return 0; }  // older name: 1 = fused, 0 = default
  if (!strcmp(name, "attn_variant")) { fb::g_flag_attn = value; return 0; }
  if (!strcmp(name, "pdl")) { fb::g_flag_pdl = value ? 1 : 0; return 0; }
  if (!strcmp(name, "gemm_cl4")) { fb::g_flag_cl4 = value ? 1 : 0; return 0; }
  if (!strcmp(name, "step_graph")) { fb::g_flag_graph = value ? 1 : 0; return 0; }
  if (!strcmp(name, "gemm_big")) { fb::g_flag_big = value; return 0; }
  if (!strcmp(name, "dequant_overlap")) { fb::g_flag_dqo = value ? 1 : 0; return 0; }
  if (!strcmp(name, "ln_reread")) { fb::g_flag_lnr = value ? 1 : 0; return 0; }
  return fb::fail(std::string("set_flag: unknown flag ") + name);
}
void fluxb200_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(fb::g_prof_mu);
  __atomic_store_n(&fb::g_prof_on, on != 0, __ATOMIC_RELAXED);
}
// Synchronises the device, folds every recorded event pair into per-kind totals and clears the records.
// Arrays must hold fluxb200_profile_kinds() entries. ms = summed kernel time, count = event pairs.
int fluxb200_profile_collect(double* ms, double* flops, double* bytes, unsigned long long* count) {
  using namespace fb;
  if (cudaDeviceSynchronize() != cudaSuccess) return fail("profile_collect: device synchronize failed");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int k = 0; k < KK_COUNT; ++k) ms[k] = flops[k] = bytes[k] = 0, count[k] = 0;
  for (auto& r : g_prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.kind] += t, flops[r.kind] += r.flops, bytes[r.kind] += r.bytes, count[r.kind] += 1;
    }
    g_prof_pool.push_back({r.a, r.b});
  }
  g_prof_recs.clear();
  return 0;
}
int fluxb200_profile_kinds(void) { return fb::KK_COUNT; }
const char* fluxb200_profile_kind_name(int k) {
  static const char* names[] = {"gemm_tcgen05", "attention_tcgen05", "ln_modulate", "qknorm_rope", "dequant",
                                "groupnorm", "misc"};
  return (k >= 0 && k < fb::KK_COUNT) ? names[k] : "?";
}
// Kernel launches issued by this library since load (all kinds / one kind).
unsigned long long fluxb200_launch_count(int kind) {
  unsigned long long t = 0;
  for (int k = 0; k < fb::KK_COUNT; ++k)
    if (kind < 0 || kind == k) t += __atomic_load_n(&fb::g_launches[k], __ATOMIC_RELAXED);
  return t;
}
}
