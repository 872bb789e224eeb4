// Shared pieces of the warp-per-element tuned kernels: the bank-conflict-free node swizzle and the
// TMA (cp.async.bulk) / mbarrier wrappers (PTX ISA: cp.async.bulk, mbarrier).
#pragma once
#include <cstdint>

#include "launch.cuh"

namespace tb {

TB_DEV int swz_pos(int n) {
    const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
    return (k << 4) | (((j ^ k) & 3) << 2) | ((i ^ k) & 3);
}

// ---- TMA / mbarrier helpers (PTX ISA: cp.async.bulk, mbarrier) ---------------------------------------
TB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
TB_DEV void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
TB_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~1 us passes)
// instead of spinning through the issue slots of the other resident warps (the kernels are issue- and power-bound;
// ncu counted 8% extra warp instructions from the plain polling loop)
TB_DEV bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000u)
        : "memory");
    return ok != 0;
}
TB_DEV void tma_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// L2 eviction priorities for streaming tiles (createpolicy + .L2::cache_hint): data that is read once should leave
// L2 first, data that the same CTA touches again a few microseconds later should stay
TB_DEV uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
TB_DEV uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
TB_DEV void tma_load_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
TB_DEV void tma_store_hint(void *dst, uint32_t src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src),
                 "r"(bytes), "l"(policy)
                 : "memory");
}
TB_DEV void tma_reduce_add_f64_hint(void *dst, uint32_t src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.L2::cache_hint.add.f64 [%0], [%1], %2, %3;" ::"l"(dst),
                 "r"(src), "r"(bytes), "l"(policy)
                 : "memory");
}
TB_DEV void tma_store(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}
// global[i] += shared[i] (FP64, round to nearest) performed by the L2's reduction units: UBLKRED.ADD.F64 in SASS
TB_DEV void tma_reduce_add_f64(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src),
                 "r"(bytes)
                 : "memory");
}
TB_DEV void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
TB_DEV void tma_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
TB_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace tb
