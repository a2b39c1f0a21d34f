// emu.cpp — TEST INFRASTRUCTURE ONLY: fiber scheduler of the SIMT emulation declared in
// tests/emu/cuda_runtime.h (see the header for what this is and is not).
#include <sys/mman.h>

#include <mutex>
#include <vector>

#include "cuda_runtime.h"

namespace fgemu {

Fiber* g_cur = nullptr;
Cta g_cta;
unsigned long long g_progress = 0;

extern "C" void fgemu_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .globl fgemu_switch
    .type fgemu_switch,@function
fgemu_switch:
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
    .size fgemu_switch,.-fgemu_switch
)");

static constexpr size_t STACK = 256 << 10;
static constexpr unsigned MAX_THREADS = 1024;
static void* g_sched_sp = nullptr;
static const std::function<void()>* g_body = nullptr;
static char* g_stacks = nullptr;
static std::vector<Fiber> g_fibers;
static std::recursive_mutex g_launch_mu;

static unsigned long long g_seed = 0, g_rng = 0;
static unsigned next_rand() {
    g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull;
    return (unsigned)(g_rng >> 33);
}
// an odd stride is coprime with the (power-of-two-multiple-of-32) block sizes used here only if it shares
// no factor with them: search for one
static unsigned pick_stride(unsigned block) {
    for (;;) {
        unsigned s = (next_rand() % block) | 1u;
        unsigned a = s, b = block;
        while (b) { unsigned t = a % b; a = b; b = t; }
        if (a == 1) return s;
    }
}

void die(const char* what) {
    fprintf(stderr, "fgemu: %s (block %u, thread %u)\n", what, g_cta.bid.x, g_cur ? g_cur->tid.x : 0u);
    abort();
}

void yield() { fgemu_switch(&g_cur->sp, g_sched_sp); }

static void fiber_main() {
    (*g_body)();
    g_cur->done = true;
    g_progress++;
    fgemu_switch(&g_cur->sp, g_sched_sp);
    die("resumed a finished fiber");
}

const char* g_kernel_name = "?";
void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body) {
    std::lock_guard<std::recursive_mutex> lock(g_launch_mu);
    if (g_cur) die("nested launch");
    static const bool g_profile = getenv("FGEMU_PROFILE") != nullptr;
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (block == 0 || block > MAX_THREADS || (block & 31)) die("block size must be a multiple of 32, at most 1024");
    if (!g_stacks) {
        if (const char* e = getenv("FGEMU_SEED")) { g_seed = strtoull(e, nullptr, 10); g_rng = g_seed * 0x9E3779B97F4A7C15ull + 1; }
        g_stacks = (char*)mmap(nullptr, STACK * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == (char*)MAP_FAILED) die("mmap of the fiber stacks failed");
    }
    g_fibers.resize(block);
    // one extra 16 bytes: the kernels read 128-bit words at 16-byte aligned offsets up to the end
    unsigned char* dyn = (unsigned char*)aligned_alloc(128, ((smem + 127) & ~(size_t)127) + 128);
    g_body = &body;
    for (unsigned b = 0; b < grid; b++) {
        memset(&g_cta, 0, sizeof(g_cta));
        g_cta.bid = {b, 0, 0};
        g_cta.bdim = {block, 1, 1};
        g_cta.gdim = {grid, 1, 1};
        g_cta.live = block;
        g_cta.dyn = dyn;
        memset(dyn, 0xCD, smem);  // shared memory is uninitialised: make reliance on zeroes visible
        for (unsigned t = 0; t < block; t++) {
            Fiber& f = g_fibers[t];
            f.tid = {t, 0, 0};
            f.lane = (int)(t & 31);
            f.warp = (int)(t >> 5);
            f.done = false;
            void** top = (void**)(g_stacks + STACK * (t + 1));
            // layout popped by fgemu_switch: r15 r14 r13 r12 rbx rbp, then `ret` into fiber_main with
            // rsp = 8 (mod 16), as after a call
            top -= 8;
            for (int i = 0; i < 6; i++) top[i] = nullptr;
            top[6] = (void*)&fiber_main;
            top[7] = nullptr;
            f.sp = top;
        }
        unsigned live = block;
        while (live) {
            const unsigned long long before = g_progress;
            // FGEMU_SEED=n: a different (pseudo-random, reproducible) thread order in every scheduler pass,
            // to shake out code that only works in lock step. Default: ascending thread ids.
            const unsigned stride = g_seed ? pick_stride(block) : 1u, first = g_seed ? next_rand() % block : 0u;
            for (unsigned i = 0; i < block; i++) {
                const unsigned t = (first + (unsigned long long)i * stride) % block;
                Fiber& f = g_fibers[t];
                if (f.done) continue;
                g_cur = &f;
                fgemu_switch(&g_sched_sp, f.sp);
                if (f.done) {
                    live--;
                    g_cta.live = live;
                    // threads that exit do not take part in later barriers: release one that only waited for them
                    if (live && g_cta.bar_arrived >= live) {
                        g_cta.bar_arrived = 0;
                        g_cta.orv[(g_cta.bar_gen + 1) & 1] = 0;
                        g_cta.bar_gen++;
                        g_progress++;
                    }
                }
            }
            g_cur = nullptr;
            if (live && g_progress == before) {
                fprintf(stderr, "fgemu: deadlock in kernel %s, block %u: %u threads wait for a barrier or a warp collective that cannot complete\n", g_kernel_name, b, live);
                for (unsigned w = 0; w < block / 32; w++)
                    fprintf(stderr, "  warp %u: arrived %08x drained %08x draining %d\n", w, g_cta.warps[w].arrived, g_cta.warps[w].drained, g_cta.warps[w].draining);
                fprintf(stderr, "  barrier: %u of %u arrived\n", g_cta.bar_arrived, g_cta.live);
                abort();
            }
        }
    }
    g_body = nullptr;
    free(dyn);
    if (g_profile) {  // FGEMU_PROFILE=1: host time per launch, a rough proxy for executed thread-instructions
        timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[fgemu] launch grid %u block %u smem %zu: %.1f ms\n", grid, block, smem,
                (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
    }
}

}  // namespace fgemu
