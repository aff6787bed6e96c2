// EVIDENCE FILE (not compiled, not part of the product): the "probe all, then stream" shape of the lockstep search that was measured in
// round 2 (profiles/r02/probe_lockstep_variants_v2_and_occupancy.log: 0.269 ms vs 0.201 ms for the product's shape at G = 4, PB = 2;
// bit-identical results through the emulated library and tests/test_gpu_parity.py / test_gpu_edge.py / test_gpu_golden.py on the B200).
// It lived in lidar_imu_init_b200/csrc/knn_kernels.cuh behind -DLI_KNN_V2=1 for one measurement session and was removed when it lost.
//
// Idea: the product's shape alternates {enumerate G bricks, probe them, scan what was found}; a 27-brick closing shell is seven dependent
// {hash probe -> slab loads} round trips. Here a shell is cut differently:
//   A  every lane evaluates LI_KNN_PB bricks per round and issues their hash probes back to back; what is found is appended to the
//      group's list in shared memory through a group ballot;
//   B  the group streams the CONCATENATION of the listed slabs: one cursor per group walks brick after brick in G-point steps, LI_KNN_U
//      loads in flight per lane across brick boundaries (no per-brick prologue / epilogue; trip count = max over groups of TOTAL candidates).
// Why it lost: the cursor arithmetic per candidate (advance, brick switch, id bookkeeping: ~7 instructions) costs more than the per-brick
// restarts it removes, and the kernel is instruction-issue bound.

#define LI_KNN_PB 2          // bricks evaluated (hash probes in flight) per lane and round
#define LI_KNN_LIST 32       // found bricks a group lists before it streams them

__device__ __forceinline__ uint4 li_brick_probe_issue(const uint4* __restrict__ ent, unsigned mask, unsigned long long key, unsigned& h) {
    h = li_hash(key) & mask;
    return __ldg(&ent[h]);
}
__device__ __forceinline__ bool li_brick_probe_finish(const uint4* __restrict__ ent, unsigned mask, unsigned long long key, unsigned h, uint4 e,
                                                      unsigned& first, unsigned& count) {
    for (unsigned i = 0; i <= mask; i++) {
        const unsigned long long k = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
        if (k == key) { first = e.z; count = e.w; return true; }
        if (k == LI_EMPTY_KEY) return false;
        h = (h + 1) & mask;
        e = __ldg(&ent[h]);
    }
    return false;
}

// Phase B: stream the group's listed slabs (nl entries {first, count}, group-uniform) as one candidate sequence.
template <int G, int U = LI_KNN_U>
__device__ __forceinline__ void group_stream_list(const float4* __restrict__ pool, const uint2* __restrict__ glist, int nl, float qx, float qy,
                                                  float qz, float thr, float (&ld)[5], int (&li)[5], int gl) {
    const float cap5 = __uint_as_float(0x40a00001u);
    const float thr5 = fminf(thr, cap5);
    float tau = fminf(thr5, ld[4]);
    const unsigned long long qxy = li_pack_f32x2(qx, qy);
    int b = 0, o = 0;            // the group's cursor: brick b of the list, offset o inside its slab
    unsigned f = 0, c = 0;
    if (nl > 0) { const uint2 e = glist[0]; f = e.x; c = e.y; }
    float4 a[U];
    int id[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {   // prologue: U loads in flight
        ok[u] = b < nl && (unsigned)(o + gl) < c;
        id[u] = (int)f + o + gl;
        if (ok[u]) a[u] = __ldg(pool + id[u]);
        o += G;
        if (b < nl && (unsigned)o >= c) { b++; o = 0; if (b < nl) { const uint2 e = glist[b]; f = e.x; c = e.y; } }
    }
    bool work = b < nl;
#pragma unroll
    for (int u = 0; u < U; u++) work = work || ok[u];
    while (__any_sync(LI_FULL, work)) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (ok[u]) {
                const float d = li_dist2_packed(qxy, qz, a[u]);
                if (d < tau) { local_insert(ld, li, d, id[u]); tau = fminf(thr5, ld[4]); }
            }
            ok[u] = b < nl && (unsigned)(o + gl) < c;            // refill this buffer from the cursor
            id[u] = (int)f + o + gl;
            if (ok[u]) a[u] = __ldg(pool + id[u]);
            o += G;
            if (b < nl && (unsigned)o >= c) { b++; o = 0; if (b < nl) { const uint2 e = glist[b]; f = e.x; c = e.y; } }
        }
        work = b < nl;
#pragma unroll
        for (int u = 0; u < U; u++) work = work || ok[u];
    }
}

// Shell logic identical to knn5_lockstep; only the body of a shell differs:
//
//     int nl = 0;   // bricks in the group's list (group-uniform)
//     for (int base = 0; __any_sync(LI_FULL, base < total); base += PB * G) {
//         // phase A: PB bricks per lane -- box distance, shell / bound tests, then ALL probes issued (li_brick_probe_issue) ...
//         // ... and only then examined (li_brick_probe_finish); found bricks appended through a group ballot:
//         //     const unsigned fm = grp_ballot<G>(found, gbase);
//         //     if (found) glist[nl + __popc(fm & ltg)] = make_uint2(first, count);
//         //     nl += __popc(fm);
//         if (__any_sync(LI_FULL, nl > LI_KNN_LIST - PB * G)) { group_stream_list<G>(...); nl = 0; }   // a list could overflow
//     }
//     group_stream_list<G>(M.pool, glist, nl, qx, qy, qz, thr, ld, li, gl);
//     group_merge<G>(ld, li, gd, gi, gl, gbase);
//
// glist = __shared__ uint2 [warps][32 / G][LI_KNN_LIST]: 8 KB per 128-thread block at G = 4.
