// TEST INFRASTRUCTURE ONLY — fiber scheduler behind tests/emu/cuda_emu.h.
// One OS thread; each CUDA thread of the running block is a fiber with its own
// stack.  A fiber runs until it reaches a block/warp barrier or returns; the
// scheduler sweeps the fibers round-robin and releases a barrier once every
// live participant has arrived.
#include "cuda_emu.h"
#include <cstdio>
#include <sys/mman.h>

namespace cs_emu {

dim3 g_gridDim, g_blockDim;
uint3 g_blockIdx, g_threadIdx;
unsigned char* g_dyn_smem = nullptr;
int g_sync_or = 0;
int g_sync_count = 0;

namespace {
enum State { READY, AT_BLOCK, AT_WARP, DONE };
struct Fiber {
    void* sp;
    State st;
    uint3 tid;
};
constexpr size_t kStack = 96 * 1024;
std::vector<Fiber> fibers;
std::vector<void*> stacks;
void* sched_sp = nullptr;
int cur = -1;
const std::function<void()>* cur_body = nullptr;
std::vector<unsigned long long> slots;  // 32 x u64 per warp

extern "C" void cs_emu_swap(void** from_sp, void* to_sp);
asm(R"(
.text
.globl cs_emu_swap
.type cs_emu_swap,@function
cs_emu_swap:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cs_emu_swap,.-cs_emu_swap
)");

void yield_to_sched() { cs_emu_swap(&fibers[cur].sp, sched_sp); }

extern "C" void cs_emu_entry() {
    (*cur_body)();
    fibers[cur].st = DONE;
    yield_to_sched();
    abort();  // never resumed
}

void* stack_for(size_t i) {
    while (stacks.size() <= i) {
        void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) { perror("mmap"); abort(); }
        stacks.push_back(p);
    }
    return stacks[i];
}

void prime(Fiber& f, size_t i) {
    // Lay the stack out so that the first cs_emu_swap() into it pops six zeroed
    // callee-saved registers and `ret`s into cs_emu_entry with rsp%16 == 8, as
    // the SysV ABI expects at function entry.
    uintptr_t top = ((uintptr_t)stack_for(i) + kStack) & ~(uintptr_t)15;
    uint64_t* sp = (uint64_t*)top;
    *--sp = 0;                          // fake return address of cs_emu_entry (never used)
    *--sp = (uint64_t)&cs_emu_entry;    // popped by `ret`
    for (int k = 0; k < 6; ++k) *--sp = 0;
    f.sp = sp;
}
}  // namespace

unsigned lane_id() {
    unsigned lin = g_threadIdx.x + g_blockDim.x * (g_threadIdx.y + g_blockDim.y * g_threadIdx.z);
    return lin & 31u;
}
static unsigned warp_id_of(int i) { return (unsigned)i >> 5; }
unsigned long long* warp_slots64() { return &slots[warp_id_of(cur) * 32]; }
unsigned* warp_slots() { return (unsigned*)&slots[warp_id_of(cur) * 32]; }  // lane i -> s[i] (low half of slots; separate use)

void block_barrier() { fibers[cur].st = AT_BLOCK; yield_to_sched(); }
void warp_barrier() { fibers[cur].st = AT_WARP; yield_to_sched(); }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const size_t nt = (size_t)block.x * block.y * block.z;
    if (nt == 0 || (size_t)grid.x * grid.y * grid.z == 0) return;
    g_gridDim = grid; g_blockDim = block;
    std::vector<unsigned char> dyn(smem + 64);
    g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    cur_body = &body;
    fibers.assign(nt, Fiber{});
    slots.assign(((nt + 31) / 32) * 32, 0ull);
    const size_t nwarps = (nt + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = {bx, by, bz};
        for (size_t i = 0; i < nt; ++i) {
            Fiber& f = fibers[i];
            f.st = READY;
            f.tid = {(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / ((size_t)block.x * block.y))};
            prime(f, i);
        }
        size_t done = 0;
        while (done < nt) {
            bool ran = false;
            for (size_t i = 0; i < nt; ++i) {
                if (fibers[i].st != READY) continue;
                ran = true;
                cur = (int)i;
                g_threadIdx = fibers[i].tid;
                cs_emu_swap(&sched_sp, fibers[i].sp);
                if (fibers[i].st == DONE) ++done;
            }
            // release barriers whose live participants have all arrived
            size_t live = 0, at_block = 0;
            for (size_t i = 0; i < nt; ++i) { if (fibers[i].st != DONE) ++live; if (fibers[i].st == AT_BLOCK) ++at_block; }
            bool released = false;
            if (live && at_block == live) { for (auto& f : fibers) if (f.st == AT_BLOCK) f.st = READY; released = true; }
            for (size_t w = 0; w < nwarps; ++w) {
                size_t lo = w * 32, hi = std::min(nt, lo + 32), l = 0, a = 0;
                for (size_t i = lo; i < hi; ++i) { if (fibers[i].st != DONE) ++l; if (fibers[i].st == AT_WARP) ++a; }
                if (l && a == l) { for (size_t i = lo; i < hi; ++i) if (fibers[i].st == AT_WARP) fibers[i].st = READY; released = true; }
            }
            if (!ran && !released && done < nt) {
                fprintf(stderr, "cs_emu: barrier deadlock (divergent __syncthreads / warp collective)\n");
                abort();
            }
        }
    }
    cur = -1; cur_body = nullptr; g_dyn_smem = nullptr;
}

}  // namespace cs_emu
