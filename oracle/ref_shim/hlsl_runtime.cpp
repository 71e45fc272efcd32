// TEST INFRASTRUCTURE — runtime half of hlsl_cpu.h: the shader registry, the fiber scheduler that executes one thread group, and the
// C entry point tests call (same signature as liboracle.so's nrd_oracle_dispatch, so oracle/runner.py can drive either engine).
#include "hlsl_cpu.h"

#include <omp.h>
#include <sys/mman.h>

#include <map>

namespace hlsl {

static std::map<std::string, ShaderEntry>& registry() { static std::map<std::string, ShaderEntry> r; return r; }
void registerShader(const ShaderEntry& e) { registry()[e.identifier] = e; }

// probeMirror: per OS thread, per ( spatial pass, lobe ): [0] taps seen, [1] taps that took the "mirrored" branch; folded by nrd_refshader_probe
static constexpr int kMaxProbeThreads = 512, kProbeSlots = 6;   // slot = pass * 2 + lobe, pass 0 pre-pass / 1 blur / 2 post-blur
static uint64_t g_probe[kMaxProbeThreads][kProbeSlots][2];
static int g_probePass = 0;          // set per dispatch from the shader identifier
static bool g_probeTwoLobes = true;  // NRD_SIGNAL=BOTH: the first 8 taps of a thread are the diffuse lobe's
static bool g_probeSpecOnly = false;
bool probeMirror(bool mirrored) {
    Fiber& f = currentFiber();
    const int lobe = g_probeTwoLobes ? (int)((f.probeCalls >> 3) & 1u) : (g_probeSpecOnly ? 1 : 0);
    f.probeCalls++;
    uint64_t* c = g_probe[omp_get_thread_num() % kMaxProbeThreads][g_probePass * 2 + lobe];
    c[0]++;
    c[1] += mirrored ? 1u : 0u;
    return mirrored;
}

static thread_local GroupRun t_group;
GroupRun& groupRun() { return t_group; }

static constexpr size_t kStackBytes = 512 * 1024;

static void fiberMain() {
    GroupRun& g = t_group;
    Fiber& f = g.fibers[g.current];
    uint3 dt = g.groupID * g.groupSize + f.groupThreadID;
    g.entry(f.groupThreadID, g.groupID, dt, f.groupIndex);
    f.state = 2;   // returns to g.sched through uc_link
}

void runGroup(const ShaderModule& module, void (*entry)(uint3, uint3, uint3, uint), uint3 groupID, uint3 groupSize, bool useFibers) {
    GroupRun& g = t_group;
    for (const SmemReg& r : module.smem) memset(r.address(), 0, r.bytes);
    g.entry = entry; g.groupID = groupID; g.groupSize = groupSize; g.useFibers = useFibers;
    const int n = (int)(groupSize.x * groupSize.y * groupSize.z);
    if (!useFibers) {
        if (g.fibers.empty()) g.fibers.resize(1);
        g.current = 0;
        for (uint z = 0; z < groupSize.z; z++) for (uint y = 0; y < groupSize.y; y++) for (uint x = 0; x < groupSize.x; x++) {
            uint3 t(x, y, z);
            g.fibers[0].probeCalls = 0;
            entry(t, groupID, uint3(groupID * groupSize + t), (z * groupSize.y + y) * groupSize.x + x);
        }
        return;
    }
    if ((int)g.fibers.size() < n) g.fibers.resize(n);
    for (int i = 0; i < n; i++) {
        Fiber& f = g.fibers[i];
        if (!f.stack) {
            f.stack = (char*)mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (f.stack == MAP_FAILED) { perror("hlsl_cpu: fiber stack"); abort(); }
        }
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = kStackBytes; f.ctx.uc_link = &g.sched;
        makecontext(&f.ctx, fiberMain, 0);
        f.state = 0; f.xchgSeq = 0; f.probeCalls = 0; f.groupIndex = (uint)i;
        f.groupThreadID = uint3((uint)i % groupSize.x, ((uint)i / groupSize.x) % groupSize.y, (uint)i / (groupSize.x * groupSize.y));
    }
    for (;;) {
        // one round: every runnable fiber advances to its next yield ( barrier, quad exchange or end )
        for (int i = 0; i < n; i++) if (g.fibers[i].state == 0) { g.current = i; swapcontext(&g.sched, &g.fibers[i].ctx); }
        int runnable = 0, waiting = 0;
        for (int i = 0; i < n; i++) { runnable += g.fibers[i].state == 0; waiting += g.fibers[i].state == 1; }
        if (runnable) continue;
        if (!waiting) break;
        for (int i = 0; i < n; i++) if (g.fibers[i].state == 1) g.fibers[i].state = 0;   // everybody alive reached the barrier
    }
    g.current = -1;
}
}  // namespace hlsl

using namespace hlsl;

extern "C" {
// returns 0 ok, 1 unknown shader, 2 bad arguments ( binding count or constant-buffer size do not fit the shader's declarations )
__attribute__((visibility("default"))) int nrd_refshader_dispatch(const char* shaderIdentifier, const void* constants, uint32_t constantsSize, const HostTexture* textures,
                                                                   uint32_t texturesNum, uint32_t gridW, uint32_t gridH, uint32_t /*flags*/) {
    auto it = registry().find(shaderIdentifier ? shaderIdentifier : "");
    if (it == registry().end()) return 1;
    const ShaderEntry& e = it->second;
    ShaderModule& m = *e.module;
    if (m.srv.size() + m.uav.size() != texturesNum) { fprintf(stderr, "refshader %s: %zu SRV + %zu UAV declared, %u bound\n", e.identifier, m.srv.size(), m.uav.size(), texturesNum); return 2; }
    for (uint32_t i = 0; i < texturesNum; i++) {
        TexData* t = i < m.srv.size() ? m.srv[i] : m.uav[i - m.srv.size()];
        if (!t) { fprintf(stderr, "refshader %s: register hole at binding %u\n", e.identifier, i); return 2; }
        t->data = (uint8_t*)textures[i].data; t->w = (int)textures[i].width; t->h = (int)textures[i].height; t->pitch = (int)textures[i].pitchBytes; t->fmt = textures[i].format;
    }
    if (!m.constants.empty() && constantsSize && !loadConstants(m, constants, constantsSize)) { fprintf(stderr, "refshader %s: constant buffer of %u bytes does not match the declared layout\n", e.identifier, constantsSize); return 2; }
    const std::string ident = shaderIdentifier;
    g_probePass = ident.find("REBLUR_PostBlur") == 0 ? 2 : (ident.find("REBLUR_Blur") == 0 ? 1 : 0);
    g_probeTwoLobes = ident.find("NRD_SIGNAL=BOTH") != std::string::npos;
    g_probeSpecOnly = ident.find("NRD_SIGNAL=SPEC") != std::string::npos;
    const uint3 gs(e.gx, e.gy, e.gz);
    const long groups = (long)gridW * (long)gridH;
#pragma omp parallel for schedule(dynamic, 1)
    for (long k = 0; k < groups; k++) runGroup(m, e.entry, uint3((uint)(k % gridW), (uint)(k / gridW), 0u), gs, e.useFibers);
    return 0;
}
__attribute__((visibility("default"))) int nrd_refshader_count() { return (int)registry().size(); }
__attribute__((visibility("default"))) const char* nrd_refshader_name(int i) { for (auto& kv : registry()) if (i-- == 0) return kv.second.identifier; return nullptr; }
// out[0] = taps whose mirror predicate was evaluated since the last reset, out[1] = how many of them were "mirrored"; then the same pair per
// ( pass, lobe ) slot: out[2 + 2 * slot], out[3 + 2 * slot], slot = pass * 2 + lobe ( 14 values in all )
__attribute__((visibility("default"))) void nrd_refshader_probe(uint64_t* out, int reset) {
    uint64_t v[2 + 2 * kProbeSlots] = {};
    for (int i = 0; i < kMaxProbeThreads; i++)
        for (int k = 0; k < kProbeSlots; k++) {
            v[0] += g_probe[i][k][0]; v[1] += g_probe[i][k][1];
            v[2 + 2 * k] += g_probe[i][k][0]; v[3 + 2 * k] += g_probe[i][k][1];
            if (reset) g_probe[i][k][0] = g_probe[i][k][1] = 0;
        }
    if (out) memcpy(out, v, sizeof(v));
}
__attribute__((visibility("default"))) void nrd_refshader_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
}
