// particlesolver_b200/csrc/ps_fluid_lists.cuh — the neighbour lists K6 (PBF lambda) leaves for K7 (PBF delta-p) and the other
// passes over the same neighbourhoods (XSPH / vorticity / density diagnostics).
//
// NOT the reference's lists (500 slots = 2 KB per particle, 4x over-allocated, strided by thread, integration.cu:70).
// A warp's 32 lists are interleaved in a region of `list_rows` rows of 128 bytes (PsParams.neighbor_list_rows, default 512 >= the
// 500-neighbour cap, so a list always fits): row k of warp w is the line pool[(w * list_rows + k) * 32 ..], entry `lane` of it the
// k-th accepted neighbour of that lane's particle, in the reference's traversal order.  A lane's list is its column: K6 appends
// to it straight from the walk (no queue, no padding — rows past a lane's neighbour count are never written and never read, and
// an untouched row costs address space, not memory traffic), K7 reads the rows in lock-step, fully coalesced; the count is
// num_neighbors.  No allocator: the region of a warp is a function of its index.
// Per warp one status word (ps_* array PS_ARR_NEIGHBOR_ROWS): 0 = the warp holds no active fluid particle, 1 = list valid, or
//   kListOverflow: the warp has no list (neighbor_list_rows too small for its longest list) and K7 walks the grid for it.
// The results do not depend on which path a warp takes (tests/test_gpu_parity.py::test_neighbour_list_paths_agree_bit_for_bit).
#pragma once
#include "ps_common.cuh"

constexpr u32 kListRecord = PS_LIST_RECORD_WORDS;
constexpr u32 kListOverflow = 0xffffffffu;

#ifndef PS_LIST_READ_ROWS
#define PS_LIST_READ_ROWS 6  // list rows a reader fetches per trip = position gathers in flight per lane
#endif

// One neighbour's terms of the lambda pass (reference collideCellRadius / findLambdasD, integration_kernel.cuh:521-593) with explicit
// roundings, in the association of the staged kernel's PTX body (PS_K6_VISIT): every K6 variant produces the same bits.
//   ro += (H^2 - r^2)^3;  c = cs (H - r)^2 / r  (cs = -SPIKY / rho0; spikyGrad / rho0 = r_vec * c);  g += r_vec * c;  denom += c^2 r^2
__device__ __forceinline__ void ps_lambda_terms(float rx, float ry, float rz, float cs, float &ro, float &gx, float &gy, float &gz, float &denom) {
    const float r2 = __fmaf_rn(rz, rz, __fmaf_rn(rx, rx, __fmul_rn(ry, ry)));
    const float ir = rsqrtf(r2), rl = __fmul_rn(r2, ir), h2 = __fsub_rn(PS_H2, r2);
    ro = __fmaf_rn(h2, __fmul_rn(h2, h2), ro);
    if (rl >= 0.0001f) {  // coincident particles contribute no gradient (false for NaN = r2 == 0 as well)
        const float hm = __fsub_rn(PS_H, rl);
        const float c = __fmul_rn(ir, __fmul_rn(hm, __fmul_rn(hm, cs)));
        gx = __fmaf_rn(rx, c, gx); gy = __fmaf_rn(ry, c, gy); gz = __fmaf_rn(rz, c, gz);
        denom = __fmaf_rn(r2, __fmul_rn(c, c), denom);
    }
}
// lambda_i = -(rho_i / rho0 - 1) / (sum_k |grad_k C|^2 + RELAX), rho_i including the self term poly6(0) = POLY6 H^6 (:589)
__device__ __forceinline__ float ps_lambda_from_sums(float ro, float denom, float gx, float gy, float gz, float inv_w, float inv_ro0) {
    const float rho = __fmul_rn(__fadd_rn(ro, PS_H6), __fmul_rn(PS_POLY6, inv_w));
    const float g2 = __fmaf_rn(gz, gz, __fmaf_rn(gy, gy, __fmul_rn(gx, gx)));
    return -__fdividef(__fmaf_rn(rho, inv_ro0, -1.f), __fadd_rn(__fadd_rn(denom, g2), PS_RELAX));
}

// Keeps a loaded value "used" at this point of the program: without it the compiler sinks the gathers of a trip into the guarded
// blocks that consume them (seen in the SASS of round 2's K7: LDG.128, wait, compute, next LDG.128 ...), which serialises the L1 / L2
// round trips the trip was meant to overlap.
__device__ __forceinline__ void ps_pin(float4 &v) { asm volatile("" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)); }
__device__ __forceinline__ void ps_pin(float &v) { asm volatile("" : "+f"(v)); }

// f(rx, ry, rz, j) for the cnt listed neighbours j of sorted slot i, in K6's traversal order.  L: this lane's column of its warp's region.
// A trip requests U list rows, then the U positions they name (all in flight together), then runs the U bodies.
template <class F>
__device__ __forceinline__ void ps_for_each_listed(const u32 *__restrict__ L, u32 cnt, u32 i, float4 pi, const float4 *__restrict__ spos, F &&f) {
    constexpr int U = PS_LIST_READ_ROWS;
    for (u32 r = 0; r < cnt; r += U) {
        u32 j[U];
        float4 pj[U];
#pragma unroll
        for (int k = 0; k < U; k++) j[k] = (r + k < cnt) ? __ldcs(L + (size_t)(r + k) * 32) : i;
#pragma unroll
        for (int k = 0; k < U; k++) pj[k] = __ldg(spos + j[k]);
#pragma unroll
        for (int k = 0; k < U; k++) ps_pin(pj[k]);
#pragma unroll
        for (int k = 0; k < U; k++)
            if (r + k < cnt) f(pi.x - pj[k].x, pi.y - pj[k].y, pi.z - pj[k].z, j[k]);
    }
}
// the same for a body that also wants lambda_j (K7): f.apply(rx, ry, rz, lambda_j); the lambda gathers travel with the position gathers
template <class F>
__device__ __forceinline__ void ps_for_each_listed_lambda(const u32 *__restrict__ L, u32 cnt, u32 i, float4 pi, const float4 *__restrict__ spos,
                                                          const float *__restrict__ lambda, F &&f) {
    constexpr int U = PS_LIST_READ_ROWS;
    for (u32 r = 0; r < cnt; r += U) {
        u32 j[U];
        float4 pj[U];
        float lj[U];
#pragma unroll
        for (int k = 0; k < U; k++) j[k] = (r + k < cnt) ? __ldcs(L + (size_t)(r + k) * 32) : i;
#pragma unroll
        for (int k = 0; k < U; k++) { pj[k] = __ldg(spos + j[k]); lj[k] = __ldg(lambda + j[k]); }
#pragma unroll
        for (int k = 0; k < U; k++) { ps_pin(pj[k]); ps_pin(lj[k]); }
#pragma unroll
        for (int k = 0; k < U; k++)
            if (r + k < cnt) f.apply(pi.x - pj[k].x, pi.y - pj[k].y, pi.z - pj[k].z, lj[k]);
    }
}
__device__ __forceinline__ const u32 *ps_list_column(const u32 *pool, u32 warp, u32 list_rows, int lane) {
    return pool + (size_t)warp * list_rows * 32 + lane;
}

// ps_fluid_staged.cu: rows of the dump region behind the pool (writes of a warp whose allocation failed land there)
u32 ps_staged_dump_rows();
// K6 with TMA-staged neighbour rows.  spos[j].w must hold the bit pattern of j (ps_launch_reorder(..., slot_in_w = true)).
// CTAs whose geometry the staging does not cover walk the grid inside the same kernel (same results, same lists).
// dump_offset: element offset of the dump region behind the warps' regions (ps_neighbor_list_elems sizes both).
void ps_launch_find_lambdas_staged(float *lambda, u32 *num_neighbors, const float4 *spos, const float *sw, const int *sphase, const u32 *index,
                                   const u32 *cell_begin, const float *ros, u32 n, u32 n_owned, float ghost_xmin, float ghost_xmax, GridDesc g,
                                   const StencilDesc &st, bool zero_nonfluid, u32 *pool, u32 *recs, u32 list_rows, size_t dump_offset, int device,
                                   cudaStream_t s);
