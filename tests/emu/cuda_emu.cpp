// cuda_emu.cpp — fibers, rendezvous and the CTA scheduler of the SIMT emulator (TEST INFRASTRUCTURE,
// see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {

Fiber* cur = nullptr;
dim3 g_block_idx, g_block_dim, g_grid_dim;
unsigned char* g_smem = nullptr;

namespace {
void* g_sched_sp = nullptr;
const std::function<void()>* g_body = nullptr;
std::vector<Fiber> g_fibers;
std::vector<Warp> g_warps;
int g_barrier_arrived = 0, g_barrier_gen = 0, g_live = 0;
unsigned long long g_progress = 0;  // deposits, releases and exits: a scheduler round without any is a deadlock
const size_t STACK_BYTES = 256 * 1024;
}  // namespace

// switch stacks: saves the callee-saved registers of the System V ABI on the current stack, stores the
// stack pointer to *from, loads `to` and returns on the other stack
extern "C" void emu_switch(void** from, void* to);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");

static void fiber_entry()
{
    (*g_body)();
    cur->done = true;
    --g_live;
    ++g_progress;
    emu_switch(&cur->sp, g_sched_sp);
    abort();  // a finished fiber is never resumed
}

void yield() { emu_switch(&cur->sp, g_sched_sp); }

static Group& group_for(Warp& w, uint32_t mask)
{
    for (int i = 0; i < w.n_groups; ++i)
        if (w.groups[i].mask == mask) return w.groups[i];
    if (w.n_groups >= 8) {
        fprintf(stderr, "emu: more than 8 distinct collective masks in one warp\n");
        abort();
    }
    Group& g = w.groups[w.n_groups++];
    g = Group();
    g.mask = mask;
    return g;
}

void gather(uint32_t mask, uint64_t v, uint64_t out[32])
{
    Fiber* me = cur;
    const uint32_t bit = 1u << me->lane;
    if (!(mask & bit)) {
        fprintf(stderr, "emu: lane %d calls a collective whose mask %08x does not name it\n", me->lane, mask);
        abort();
    }
    Group& g = group_for(*me->warp, mask);
    while (g.leaving) yield();  // the previous collective of this group is still being read
    g.vals[me->lane] = v;
    g.arrived |= bit;
    ++g_progress;
    if (g.arrived == mask)
        g.leaving = true;
    else
        while (!g.leaving) yield();
    for (int i = 0; i < 32; ++i) out[i] = g.vals[i];
    g.left |= bit;
    ++g_progress;
    if (g.left == mask) {
        g.arrived = 0;
        g.left = 0;
        g.leaving = false;
    }
}

void cta_barrier()
{
    const int gen = g_barrier_gen;
    ++g_progress;
    if (++g_barrier_arrived == g_live) {
        g_barrier_arrived = 0;
        ++g_barrier_gen;
        return;
    }
    while (gen == g_barrier_gen) yield();
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body)
{
    const int threads = (int)(block.x * block.y * block.z);
    const int warps = (threads + 31) / 32;
    if ((int)g_fibers.size() < threads) {
        const size_t old = g_fibers.size();
        g_fibers.resize(threads);
        for (size_t i = old; i < g_fibers.size(); ++i) g_fibers[i].stack = (unsigned char*)aligned_alloc(64, STACK_BYTES);
    }
    std::vector<unsigned char> smem(smem_bytes + 64);
    g_body = &body;
    g_block_dim = block;
    g_grid_dim = grid;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_block_idx = dim3(bx, by, bz);
                g_smem = (unsigned char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
                g_warps.assign(warps, Warp());
                g_barrier_arrived = 0;
                g_live = threads;
                for (int t = 0; t < threads; ++t) {
                    Fiber& f = g_fibers[t];
                    f.tid.x = t % block.x;
                    f.tid.y = (t / block.x) % block.y;
                    f.tid.z = t / (block.x * block.y);
                    f.lane = t & 31;
                    f.warp = &g_warps[t >> 5];
                    f.done = false;
                    // initial frame: six callee-saved registers, then the entry point as the return address;
                    // after the `ret` the stack pointer is 8 mod 16, as at any function entry
                    uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
                    void** sp = (void**)top;
                    *--sp = nullptr;
                    *--sp = (void*)&fiber_entry;
                    for (int r = 0; r < 6; ++r) *--sp = nullptr;
                    f.sp = sp;
                }
                while (g_live > 0) {
                    const unsigned long long before = g_progress;
                    for (int t = 0; t < threads; ++t) {
                        if (g_fibers[t].done) continue;
                        cur = &g_fibers[t];
                        emu_switch(&g_sched_sp, cur->sp);
                    }
                    if (g_progress == before) {
                        fprintf(stderr, "emu: every live thread of CTA (%u,%u) waits and none can proceed - a collective "
                                        "whose mask names a thread that never arrives\n", bx, by);
                        abort();
                    }
                }
                cur = nullptr;
            }
    g_body = nullptr;
}

}  // namespace emu
