// Error channel and device queries of the C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace mvn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        else return 148;
    }
    return cached;
}
// Programmatic dependent launch: env MVN_PDL=0/1 pins it; otherwise the host chooses per model with mvn_set_pdl().  Measured on
// B200: a single chain of kernels (C2, and C3 where the ConvMixer chain is short) gains 3 - 4 % from dependents whose prologue runs
// under the predecessor's tail; with two long encoder chains on two streams (C4 / C5) the early-scheduled CTAs hold SM slots the
// other chain's kernels would have used, and the step is 1.5 - 3 % SLOWER -- so the model switches it off there.
static int g_pdl = -1;      // -1: not chosen yet (default on)
static int pdl_env() {
    static int v = -2;
    if (v == -2) { const char* e = getenv("MVN_PDL"); v = !e ? -1 : (e[0] == '0' ? 0 : 1); }
    return v;
}
bool pdl_enabled() {
    const int e = pdl_env();
    if (e >= 0) return e == 1;
    return g_pdl != 0;
}
extern "C" void mvn_set_pdl(int on) { g_pdl = on ? 1 : 0; }
static const uint32_t* g_step_ctr = nullptr;
const uint32_t* step_counter() { return g_step_ctr; }
static long long g_launches = 0;
void note_launch() { ++g_launches; }
static long long g_tier[TIER_N] = {0, 0, 0, 0};
void count_tier(int tier) { if (tier >= 0 && tier < TIER_N) ++g_tier[tier]; }

static unsigned g_prof_mask = 0;
struct ProfRec { cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof[PROF_NCLASS];
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t take_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
static thread_local int g_prof_depth = 0;      // a scope opened inside another one (ConvMixer stage -> its GEMMs) belongs to the outer class
ProfScope::ProfScope(int cls_, cudaStream_t st_) : cls(cls_), st(st_), stop(nullptr), on(false) {
    if ((g_prof_mask & (1u << cls)) && g_prof_depth == 0) {
        on = true;
        ++g_prof_depth;
        cudaEvent_t a = take_event();
        stop = take_event();
        cudaEventRecord(a, st);
        g_prof[cls].push_back({a, stop});
    }
}
ProfScope::~ProfScope() { if (on) { cudaEventRecord(stop, st); --g_prof_depth; } }
}  // namespace mvn

extern "C" void mvn_prof_enable(unsigned class_mask) { mvn::g_prof_mask = class_mask; }
extern "C" int mvn_prof_read(int cls, double* total_ms, long long* count) {
    using namespace mvn;
    if (cls < 0 || cls >= PROF_NCLASS || !total_ms || !count) { set_error("prof_read: bad arguments"); return MVN_E_BADARG; }
    double ms = 0.0;
    for (auto& r : g_prof[cls]) {
        cudaEventSynchronize(r.b);
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms += t;
        g_pool.push_back(r.a); g_pool.push_back(r.b);
    }
    *total_ms = ms; *count = (long long)g_prof[cls].size();
    g_prof[cls].clear();
    return 0;
}
extern "C" void mvn_set_step_counter(const uint32_t* dev_counter) { mvn::g_step_ctr = dev_counter; }
extern "C" long long mvn_launch_count(void) { return mvn::g_launches; }
extern "C" long long mvn_tier_count(int tier) { return (tier >= 0 && tier < mvn::TIER_N) ? mvn::g_tier[tier] : -1; }
extern "C" void mvn_tier_reset(void) { for (int i = 0; i < mvn::TIER_N; ++i) mvn::g_tier[i] = 0; }

extern "C" const char* mvn_last_error(void) { return mvn::g_err; }
extern "C" int mvn_abi_version(void) { return 1; }
extern "C" int mvn_num_sms(void) { return mvn::num_sms(); }
extern "C" int mvn_num_slabs(void) { return mvn::kSlabs; }
