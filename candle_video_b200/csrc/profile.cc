#include "profile.h"

#include <stdarg.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace ltxv {
namespace {
struct Rec {
    int cls;
    double flops;
    cudaEvent_t e0, e1;
};
bool g_on = false;
std::vector<Rec> g_recs;
std::mutex g_mu;
bool g_trace_on = false;
std::map<std::string, uint64_t> g_trace;
std::string g_trace_text;
}  // namespace

bool tracing_enabled() { return g_trace_on; }
void trace_begin() {
    std::lock_guard<std::mutex> lk(g_mu);
    g_trace.clear();
    g_trace_on = true;
}
const char* trace_end() {
    std::lock_guard<std::mutex> lk(g_mu);
    g_trace_on = false;
    g_trace_text.clear();
    for (auto& kv : g_trace) g_trace_text += kv.first + " " + std::to_string(kv.second) + "\n";
    g_trace.clear();
    return g_trace_text.c_str();
}
void trace_variant_slow(const char* fmt, ...) {
    char buf[160];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_trace_on) ++g_trace[buf];
}

bool profiling_enabled() { return g_on; }

void profiling_begin() {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& r : g_recs) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_recs.clear();
    g_on = true;
}

void profiling_end(uint64_t* launches, double* ms, double* flops) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_on = false;
    cudaDeviceSynchronize();
    for (int i = 0; i < PROF_NUM; ++i) {
        launches[i] = 0;
        ms[i] = 0;
        flops[i] = 0;
    }
    for (auto& r : g_recs) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
            launches[r.cls]++;
            ms[r.cls] += t;
            flops[r.cls] += r.flops;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_recs.clear();
}

ProfScope::ProfScope(int cls_, double flops_, cudaStream_t s_) : on(g_on), cls(cls_), flops(flops_), s(s_) {
    if (!on) return;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
}
ProfScope::~ProfScope() {
    if (!on) return;
    cudaEventRecord(e1, s);
    std::lock_guard<std::mutex> lk(g_mu);
    g_recs.push_back(Rec{cls, flops, e0, e1});
}

}  // namespace ltxv
