// host_prof.h -- where the host side of a tick spends its time.  Off unless SVB_HOST_PROFILE is set in the environment; then every
// SVB_PROF scope accumulates wall time and a table goes to stderr when the process ends.  A development aid (tools/README.md).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace svb {

struct HostProf {
    static constexpr int N = 32;
    const char* name[N] = {};
    double ms[N] = {};
    long calls[N] = {};
    bool on = std::getenv("SVB_HOST_PROFILE") != nullptr;
    ~HostProf() {
        if (!on) return;
        for (int i = 0; i < N; ++i)
            if (calls[i]) std::fprintf(stderr, "[svb host] %-28s %9ld calls %10.3f ms total %8.2f us/call\n", name[i], calls[i], ms[i], 1e3 * ms[i] / calls[i]);
    }
    static HostProf& get() {
        static HostProf p;
        return p;
    }
};
struct ProfScope {
    int id;
    std::chrono::steady_clock::time_point t0;
    bool on;
    ProfScope(int id, const char* label) : id(id), on(HostProf::get().on) {
        if (on) HostProf::get().name[id] = label, t0 = std::chrono::steady_clock::now();
    }
    ~ProfScope() {
        if (on) {
            HostProf& p = HostProf::get();
            p.ms[id] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            ++p.calls[id];
        }
    }
};
#define SVB_PROF_CAT2(a, b) a##b
#define SVB_PROF_CAT(a, b) SVB_PROF_CAT2(a, b)
#define SVB_PROF(id, label) ::svb::ProfScope SVB_PROF_CAT(svb_prof_scope_, __LINE__)(id, label)

}  // namespace svb
