// emul_map.h -- brick hash laid out in HOST memory by the storage rule of map_kernels.cuh (li_storage: brick and voxel id from
// the float BOX index, ikd_Tree.cpp:633,980), for the CPU checkers of the search kernels (tests/emul). TEST INFRASTRUCTURE.
// Needs common.cuh (MapDev, li_pack_key, li_hash) and cells.cuh (li_cells_refresh_brick, lc_u2f) included first.
#pragma once
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

struct Emul {
    float ds = 0.15f;
    std::unordered_map<unsigned long long, std::vector<float4>> bricks;   // slab content, in slab order
    std::vector<uint4> ent;
    std::vector<float4> pool;
    std::vector<unsigned long long> cocc;
    std::vector<unsigned short> cdir;
    std::vector<unsigned long long> sb_keys, sb_occ;
    std::vector<int> counters;
    MapDev M{};
    int hash_log2 = 16;
};

// li_box_index (map_kernels.cuh), plain arithmetic (compiled with -ffp-contract=off)
bool box_index(float x, float ds, int c, int& b) {
    float cf = (float)c;
    float mn = cf * ds, mx = mn + ds;
    if (x >= mn && x < mx) { b = c; return true; }
    if (x < mn) {
        float m1 = (cf - 1.0f) * ds, x1 = m1 + ds;
        if (x >= m1 && x < x1) { b = c - 1; return true; }
    } else {
        float m1 = (cf + 1.0f) * ds, x1 = m1 + ds;
        if (x >= m1 && x < x1) { b = c + 1; return true; }
    }
    b = c;
    return false;
}

void storage(float ds, float4& p, unsigned long long& key) {
    int cx = (int)floorf(p.x / ds), cy = (int)floorf(p.y / ds), cz = (int)floorf(p.z / ds);
    int bx, by, bz;
    bool ok = box_index(p.x, ds, cx, bx);
    ok = box_index(p.y, ds, cy, by) && ok;
    ok = box_index(p.z, ds, cz, bz) && ok;
    unsigned vib;
    if (ok) {
        key = li_pack_key(bx >> 3, by >> 3, bz >> 3);
        vib = (unsigned)(((bx & 7) << 6) | ((by & 7) << 3) | (bz & 7));
    } else {
        key = li_pack_key(cx >> 3, cy >> 3, cz >> 3);
        vib = LI_NO_BOX_W;
    }
    p.w = lc_u2f(vib);
}

void layout(Emul* E, bool refresh) {
    const size_t slots = (size_t)1 << E->hash_log2;
    E->ent.assign(slots, make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u));
    E->cocc.assign(slots, 0xdeadbeefdeadbeefull);   // garbage where nothing was refreshed, as on the device
    E->cdir.assign(slots * 64, (unsigned short)0xabcd);
    size_t total = 0;
    for (auto& kv : E->bricks) total += ((kv.second.size() + 7) & ~size_t(7)) + 8;
    E->pool.assign(total + 8, make_float4(NAN, NAN, NAN, 0.f));
    E->M.ent = E->ent.data();
    E->M.aux = nullptr;
    E->M.mask = (unsigned)slots - 1;
    E->M.pool = E->pool.data();
    E->M.pool_cap = E->pool.size();
    E->M.ds = E->ds;
    E->M.bshift = 3;
    E->M.cocc = E->cocc.data();
    E->M.cdir = E->cdir.data();
    // super-brick table, as li_sb_mark fills it when bricks are created
    E->sb_keys.assign(slots, LI_EMPTY_KEY);
    E->sb_occ.assign(slots, 0ull);
    E->M.sb_keys = E->sb_keys.data();
    E->M.sb_occ = E->sb_occ.data();
    E->M.sb_mask = (unsigned)slots - 1;
    for (auto& kv : E->bricks) {
        const unsigned long long bk = kv.first;
        const int kx = (int)(unsigned)(bk >> 42) - LI_CELL_LIMIT, ky = (int)((unsigned)(bk >> 21) & 0x1fffffu) - LI_CELL_LIMIT,
                  kz = (int)((unsigned)bk & 0x1fffffu) - LI_CELL_LIMIT;
        const unsigned long long key = li_pack_key(kx >> 2, ky >> 2, kz >> 2);
        unsigned h = li_hash(key) & E->M.sb_mask;
        while (E->sb_keys[h] != LI_EMPTY_KEY && E->sb_keys[h] != key) h = (h + 1) & E->M.sb_mask;
        E->sb_keys[h] = key;
        E->sb_occ[h] |= 1ull << (((kx & 3) << 4) | ((ky & 3) << 2) | (kz & 3));
    }
    size_t off = 0;
    for (auto& kv : E->bricks) {
        unsigned h = li_hash(kv.first) & E->M.mask;
        while (!(E->ent[h].x == 0xffffffffu && E->ent[h].y == 0xffffffffu)) h = (h + 1) & E->M.mask;
        E->ent[h] = make_uint4((unsigned)kv.first, (unsigned)(kv.first >> 32), (unsigned)off, (unsigned)kv.second.size());
        std::memcpy(&E->pool[off], kv.second.data(), kv.second.size() * sizeof(float4));
        if (refresh) {
            li_cells_refresh_brick(E->M, h);
            std::memcpy(kv.second.data(), &E->pool[off], kv.second.size() * sizeof(float4));   // the slab order is now the sorted one
        }
        off += ((kv.second.size() + 7) & ~size_t(7)) + 8;
    }
}

}  // namespace
