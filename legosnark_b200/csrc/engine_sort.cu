// engine_sort.cu — orchestration of the radix-partition bucket sort (sort_kernels.cuh).  Group-independent:
// the sort sees scalars and bucket ids only, so it lives in its own translation unit and both groups call it.
#include "engine_common.hpp"
#include "sort_kernels.cuh"

namespace b200 {
namespace eng {

// partition geometry for NB buckets, or false when the bucket set is outside what the shared-memory passes cover
bool partition_geometry(const MsmGeom &g, size_t n, int sms, SortGeom *out, uint32_t align_log)
{
    uint32_t lg = 0;
    while (((uint64_t)1 << lg) < g.NB) lg++;
    uint32_t low = lg > 11 ? lg - 11 : 0;  // ~2048 partitions
    if (low > PART_MAX_LOW) low = PART_MAX_LOW;
    const uint64_t NP = ((uint64_t)g.NB + ((1u << low) - 1)) >> low;
    if (NP > PART_MAX_NP || g.W > PART_STAGE_ITEMS / 32 || g.L + 1 > 2048) return false;
    SortGeom sg;
    sg.align_log = align_log;
    sg.low_bits = low;
    sg.NP = (uint32_t)NP;
    sg.tile = std::min<uint32_t>(SCAT_THREADS, PART_STAGE_ITEMS / g.W);
    size_t per = (n + (size_t)sms * 4 - 1) / ((size_t)sms * 4);
    per = (per + PART_THREADS - 1) / PART_THREADS * PART_THREADS;
    sg.hist_per_block = (uint32_t)std::min<size_t>(std::max<size_t>(per, PART_THREADS), 1024);
    *out = sg;
    return true;
}

static void set_sort_attributes(Device &D)
{
    if (D.sort_attr_set) return;
    CK(cudaFuncSetAttribute(k_part_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(part_npad(PART_MAX_NP) * 3 * 4 + PART_STAGE_ITEMS * 8)));
    CK(cudaFuncSetAttribute(k_part_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)2 << PART_MAX_LOW) * 4 + FINE_STAGE * 4 + 2049 * 4)));
    CK(cudaFuncSetAttribute(k_task_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)3 << PART_MAX_LOW) * 4 + 2 * 2049 * 4)));
    D.sort_attr_set = true;
}

// Bucket-orders the digits of n scalars.  On return (stream order) D.cnt / D.off / D.toff describe every bucket,
// D.entries holds the bucket-ordered (point, sign) words, D.meta / D.order the accumulation tasks longest first,
// D.totals = {entries, tasks, split buckets, ones, big buckets}; D.split / D.big list the multi-task buckets.
void enqueue_partition_sort(Device &D, cudaStream_t st, const MsmGeom &g, const SortGeom &sg, const uint8_t *d_flags,
                            const Fr *d_scalars, size_t n)
{
    set_sort_attributes(D);
    D.std_scalars.ensure(n * sizeof(Fr));
    const uint32_t amask = (1u << sg.align_log) - 1u, pad = amask << sg.low_bits;  // at most 2^align_log - 1 padding slots per bucket
    D.items.ensure(((size_t)g.W * n + (size_t)sg.NP * (pad + amask + 1)) * sizeof(uint2));
    D.part.ensure((size_t)PART_MAX_NP * 5 * 4);
    uint32_t *pcount = D.part.as<uint32_t>(), *pstart = pcount + PART_MAX_NP, *pcursor = pstart + PART_MAX_NP;
    uint32_t *ptasks = pcursor + PART_MAX_NP, *ptstart = ptasks + PART_MAX_NP;
    uint32_t *totals = D.totals.as<uint32_t>();
    CK(cudaMemsetAsync(pcount, 0, (size_t)sg.NP * 4, st));
    CK(cudaMemsetAsync((char *)D.totals.p + 12, 0, 4, st));  // totals[3]: scalars equal to one
    CK(cudaMemsetAsync(D.len_hist.p, 0, (size_t)(g.L + 1) * 4, st));
    uint32_t *len_hist = D.len_hist.as<uint32_t>(), *len_cursor = D.len_cursor.as<uint32_t>();
    Fr *stdsc = D.std_scalars.as<Fr>();
    LAUNCH(D, k_part_hist, cdiv(n, sg.hist_per_block), PART_THREADS, (size_t)sg.NP * 4, st, d_scalars, d_flags, n, g, sg, stdsc, pcount,
           g.ones ? D.ones_idx.as<uint32_t>() : (uint32_t *)nullptr, totals + 3);
    LAUNCH(D, k_part_scan, 1, 1024, 0, st, (const uint32_t *)pcount, sg.NP, pstart, pcursor, totals + 5, (uint32_t *)nullptr, (uint32_t *)nullptr,
           (const uint32_t *)nullptr, (uint32_t *)nullptr, 0u, pad, amask, totals + 0);  // totals[0]: entries; totals[5]: slots incl. padding
    LAUNCH(D, k_part_scatter, cdiv(n, sg.tile), SCAT_THREADS, part_scatter_smem(sg), st, (const Fr *)stdsc, n, g, sg, pcursor,
           D.items.as<uint2>());
    LAUNCH(D, k_part_sort, sg.NP, FINE_THREADS, part_sort_smem(sg, g.L), st, (const uint2 *)D.items.as<uint2>(), (const uint32_t *)pstart,
           (const uint32_t *)pcount, g, sg, D.cnt.as<uint32_t>(), D.off.as<uint32_t>(), D.entries.as<uint32_t>(), ptasks, len_hist);
    LAUNCH(D, k_part_scan, 1, 1024, 0, st, (const uint32_t *)ptasks, sg.NP, ptstart, (uint32_t *)nullptr, totals + 1, totals + 2, totals + 4,
           (const uint32_t *)len_hist, len_cursor, g.L, 0u, 0u, (uint32_t *)nullptr);
    LAUNCH(D, k_task_emit, sg.NP, FINE_THREADS, task_emit_smem(sg, g.L), st, (const uint32_t *)D.cnt.as<uint32_t>(),
           (const uint32_t *)D.off.as<uint32_t>(), (const uint32_t *)ptstart, g, sg, D.toff.as<uint32_t>(), D.meta.as<uint2>(),
           D.order.as<uint32_t>(), totals, D.split.as<uint32_t>(), D.big.as<uint32_t>(), len_cursor, D.task_bucket.as<uint32_t>());
}

}  // namespace eng
}  // namespace b200
