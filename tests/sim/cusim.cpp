// cusim.cpp -- scheduler of the CUDA execution-model emulator (see cusim.h).
#include "cusim.h"

namespace cusim {
Global g;

// x86-64 SysV context switch: save callee-saved registers + rsp.
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
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
.size cusim_switch,.-cusim_switch
)");

void yield() { cusim_switch(&g.cur->rsp, g.schedRsp); }

static void fiber_exit()
{
    Fiber* f = g.cur;
    f->done = true;
    g.alive--;
    WarpState& W = g.warps[f->lin >> 5];
    W.alive &= ~(1u << (f->lin & 31));
    // a pending __syncthreads may now be complete
    if (g.alive > 0 && g.barArrived >= g.alive) {
        g.barArrived = 0;
        g.barGen++;
    }
    cusim_switch(&f->rsp, g.schedRsp);
    abort(); // never resumed
}

static void fiber_main()
{
    (*g.body)();
    fiber_exit();
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body)
{
    const int nt = (int)(block.x * block.y * block.z);
    if (nt <= 0 || grid.x == 0 || grid.y == 0 || grid.z == 0)
        return;
    if ((int)g.fibers.size() < nt) {
        const size_t old = g.fibers.size();
        g.fibers.resize(nt);
        for (size_t i = old; i < (size_t)nt; i++)
            g.fibers[i].stack = (char*)aligned_alloc(64, g.stackSize);
    }
    g.warps.resize((nt + 31) / 32);
    g.body = &body;
    g.bDim = block;
    g.gDim = grid;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g.bIdx = dim3(bx, by, bz);
                g.alive = nt;
                g.barArrived = 0;
                for (auto& W : g.warps) {
                    W.alive = W.arrived = W.draining = W.snapMask = 0;
                }
                for (int t = 0; t < nt; t++) {
                    Fiber& f = g.fibers[t];
                    f.lin = t;
                    f.done = false;
                    f.tIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    g.warps[t >> 5].alive |= 1u << (t & 31);
                    // initial frame: 6 callee-saved regs + return address (fiber_main);
                    // keep the ABI's 16-byte alignment at function entry (rsp % 16 == 8).
                    uintptr_t top = ((uintptr_t)(f.stack + g.stackSize)) & ~(uintptr_t)15;
                    void** sp = (void**)top;
                    *(--sp) = nullptr;             // fake return address of fiber_main (alignment)
                    *(--sp) = (void*)&fiber_main;  // `ret` target
                    for (int r = 0; r < 6; r++)
                        *(--sp) = nullptr;
                    f.rsp = sp;
                }
                int remaining = nt;
                long idle = 0;
                while (remaining > 0) {
                    int progressed = 0;
                    for (int t = 0; t < nt; t++) {
                        Fiber& f = g.fibers[t];
                        if (f.done)
                            continue;
                        g.cur = &f;
                        cusim_switch(&g.schedRsp, f.rsp);
                        if (f.done) {
                            remaining--;
                            progressed++;
                        }
                    }
                    if (!progressed) {
                        if (++idle > 200000000L) {
                            fprintf(stderr, "cusim: deadlock suspected in CTA (%u,%u,%u)\n", bx, by, bz);
                            abort();
                        }
                    } else {
                        idle = 0;
                    }
                }
            }
    g.body = nullptr;
}
} // namespace cusim
