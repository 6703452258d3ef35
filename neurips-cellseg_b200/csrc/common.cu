// Error plumbing for the C-ABI: entry points return an int code and leave the
// text in a thread-local buffer read back through cs_last_error().
#include "cs_platform.cuh"
#include <cstdarg>
#include <atomic>
#include <mutex>
#include <vector>

namespace cs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }


#ifndef CS_EMU
namespace prof {
struct Rec { cudaEvent_t a, b; int slot; };
struct Slot { const char* name; double ms; unsigned long long n; };
static std::mutex g_mu;
static std::atomic<bool> g_on{false};
static std::vector<Rec> g_pending;
static std::vector<cudaEvent_t> g_free;
static std::vector<Slot> g_slots;
constexpr size_t MAX_PENDING = 1u << 16;

bool enabled() { return g_on.load(std::memory_order_relaxed); }

static void drain_locked() {
    for (Rec& r : g_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            g_slots[r.slot].ms += ms; g_slots[r.slot].n += 1;
        }
        g_free.push_back(r.a); g_free.push_back(r.b);
    }
    g_pending.clear();
}
void enable(bool on) {
    std::lock_guard<std::mutex> lk(g_mu);
    drain_locked();
    for (Slot& s : g_slots) { s.ms = 0; s.n = 0; }
    g_on.store(on);
}
static cudaEvent_t take_event() {
    if (!g_free.empty()) { cudaEvent_t e = g_free.back(); g_free.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void begin(const char* name, cudaStream_t st, int* token) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_pending.size() >= MAX_PENDING) return;
    int slot = -1;
    for (size_t i = 0; i < g_slots.size(); ++i) if (g_slots[i].name == name || !strcmp(g_slots[i].name, name)) { slot = (int)i; break; }
    if (slot < 0) { g_slots.push_back(Slot{name, 0.0, 0}); slot = (int)g_slots.size() - 1; }
    Rec r{take_event(), take_event(), slot};
    if (!r.a || !r.b) return;
    cudaEventRecord(r.a, st);
    g_pending.push_back(r);
    *token = (int)g_pending.size() - 1;
}
void end(cudaStream_t st, int token) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (token < (int)g_pending.size()) cudaEventRecord(g_pending[token].b, st);
}
int read(int index, char* name, size_t cap, double* total_ms, unsigned long long* count) {
    std::lock_guard<std::mutex> lk(g_mu);
    drain_locked();
    if (index < 0 || index >= (int)g_slots.size()) return 0;
    if (name && cap) { strncpy(name, g_slots[index].name, cap - 1); name[cap - 1] = 0; }
    if (total_ms) *total_ms = g_slots[index].ms;
    if (count) *count = g_slots[index].n;
    return 1;
}
}  // namespace prof
#endif

}  // namespace cs
