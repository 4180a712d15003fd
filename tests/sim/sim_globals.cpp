// TEST INFRASTRUCTURE ONLY: storage for the emulated CUDA built-in variables (see cuda_sim.h).
#include "cuda_sim.h"
dim3 threadIdx, blockIdx, blockDim, gridDim;

#ifdef CUSIM_FAST_SWITCH
// void cusim_switch(void **save_sp, void *load_sp): park the running fiber (callee-saved registers on its own stack, stack
// pointer into *save_sp) and continue the one whose stack pointer is load_sp.  System V x86-64.
asm(R"(
    .pushsection .text
    .globl cusim_switch
    .hidden cusim_switch
    .type cusim_switch, @function
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
    .size cusim_switch, .-cusim_switch
    .popsection
)");
#endif
