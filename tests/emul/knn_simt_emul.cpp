// knn_simt_emul.cpp -- the product's lockstep 5-NN kernel (lidar_imu_init_b200/csrc/knn_kernels.cuh: k_knn_scan, knn5_lockstep,
// group scans, merges) executed on the CPU through simt_shim.h, UNCHANGED, over a host-memory brick hash.
// TEST INFRASTRUCTURE: nothing under lidar_imu_init_b200/ loads this; the product has no CPU path.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -Wno-attributes -Wno-unknown-pragmas -I/usr/local/cuda/include knn_simt_emul.cpp
#include "simt_shim.h"
// the kernels, as the library compiles them
#include "../../lidar_imu_init_b200/csrc/knn_kernels.cuh"
#include "../../lidar_imu_init_b200/csrc/icp_kernels.cuh"
// host map (needs cells.cuh for the storage helpers; its directory is built but not used by this kernel)
#include "../../lidar_imu_init_b200/csrc/cells.cuh"
#include "emul_map.h"

namespace {
struct ScanArgs {
    MapDev M;
    ScanDev S;
    PoseD P;
    float rho2;
    int G;
};
template <int G>
void lane_entry(void* a) {
    ScanArgs* s = (ScanArgs*)a;
    k_knn_scan<G, false>(s->M, s->S, s->P, s->rho2, nullptr, 0);
}
}  // namespace

extern "C" {

void* simt_map_create(const float* xyz, int n, float ds, int hash_log2) {
    Emul* E = new Emul();
    E->ds = ds;
    E->hash_log2 = hash_log2;
    for (int i = 0; i < n; i++) {
        float4 p = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.f);
        unsigned long long key;
        storage(ds, p, key);
        E->bricks[key].push_back(p);
    }
    layout(E, true);
    return E;
}
void simt_map_destroy(void* h) { delete (Emul*)h; }

// One search pass of k_knn_scan<G> over n body points: world [n*3], neighbour coordinates [n*15], counts [n].
// pose = R[9] p[3] RLI[9] TLI[3] (row-major doubles). Returns the number of warp rendezvous executed (a cost proxy).
long long simt_knn_scan(void* h, const float* body, int n, const double* pose, float rho2, int G, float* world, float* near_xyz, int* near_cnt) {
    Emul* E = (Emul*)h;
    std::vector<float4> b(n), w(n), nv(n);
    std::vector<int> ids((size_t)n * 5, -1);
    std::vector<unsigned char> sel(n, 0);
    for (int i = 0; i < n; i++) b[i] = make_float4(body[3 * (size_t)i], body[3 * (size_t)i + 1], body[3 * (size_t)i + 2], 0.f);
    ScanArgs a;
    a.M = E->M;
    a.S.body = b.data();
    a.S.world = w.data();
    a.S.near_ids = ids.data();
    a.S.selected = sel.data();
    a.S.normvec = nv.data();
    a.S.n = n;
    memcpy(a.P.R, pose, 72);
    memcpy(a.P.p, pose + 9, 24);
    memcpy(a.P.RLI, pose + 12, 72);
    memcpy(a.P.TLI, pose + 21, 24);
    a.rho2 = rho2;
    a.G = G;
    g_b.rendezvous = 0;
    void (*fn)(void*) = G == 32 ? lane_entry<32> : G == 16 ? lane_entry<16> : G == 8 ? lane_entry<8> : G == 2 ? lane_entry<2> : lane_entry<4>;
    simt_launch(2, 64, fn, &a);   // two blocks of two warps walk the scan (the kernel's grid-stride loop over warp batches)
    for (int i = 0; i < n; i++) {
        world[3 * (size_t)i] = w[i].x; world[3 * (size_t)i + 1] = w[i].y; world[3 * (size_t)i + 2] = w[i].z;
        int cnt = 0;
        for (int k = 0; k < 5; k++) {
            const int id = ids[(size_t)i * 5 + k];
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (id >= 0) {
                q = E->pool[id];
                cnt++;
            }
            near_xyz[15 * (size_t)i + 3 * k] = q.x; near_xyz[15 * (size_t)i + 3 * k + 1] = q.y; near_xyz[15 * (size_t)i + 3 * k + 2] = q.z;
        }
        near_cnt[i] = cnt;
    }
    return (long long)g_b.rendezvous;
}

// One ICP search pass as the library runs it (k_knn_scan<G> then k_icp_plane<imu, SEARCH = true>), optionally followed by a
// reuse pass (k_icp_plane<imu, SEARCH = false>) at pose2. out160 / out160_2: [HtH 144 | Htr 12 | res_sq | m | 0 0].
// Per-point state after the LAST pass run: selected [n], normvec [n*4]; world / neighbours after the search pass.
void simt_icp_pass(void* h, const float* body, int n, const double* pose, const double* pose2, float rho2, int G, int imu_en, double* out160,
                   double* out160_2, float* world, float* near_xyz, int* near_cnt, unsigned char* selected, float* normvec) {
    Emul* E = (Emul*)h;
    std::vector<float4> b(n), w(n), nv(n, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<int> ids((size_t)n * 5, -1);
    std::vector<unsigned char> sel(n, 0);
    for (int i = 0; i < n; i++) b[i] = make_float4(body[3 * (size_t)i], body[3 * (size_t)i + 1], body[3 * (size_t)i + 2], 0.f);
    ScanArgs a;
    a.M = E->M;
    a.S.body = b.data();
    a.S.world = w.data();
    a.S.near_ids = ids.data();
    a.S.selected = sel.data();
    a.S.normvec = nv.data();
    a.S.n = n;
    auto set_pose = [&](const double* ps) {
        memcpy(a.P.R, ps, 72);
        memcpy(a.P.p, ps + 9, 24);
        memcpy(a.P.RLI, ps + 12, 72);
        memcpy(a.P.TLI, ps + 21, 24);
    };
    set_pose(pose);
    a.rho2 = rho2;
    a.G = G;
    void (*fn)(void*) = G == 32 ? lane_entry<32> : G == 16 ? lane_entry<16> : G == 8 ? lane_entry<8> : G == 2 ? lane_entry<2> : lane_entry<4>;
    simt_launch(3, 128, fn, &a);
    // the plane pass: 256-thread blocks, grid-stride (launch_plane in liinit_gpu.cu); a small grid so that several rounds run
    const unsigned grid = 3;
    std::vector<double> partials((size_t)grid * 96, 0.0);
    unsigned done = 0;
    auto plane = [&](bool search, double* out) {
        if (imu_en) {
            if (search) simt_launch_fn(grid, 256, [&]() { k_icp_plane<true, true>(a.M, a.S, a.P, partials.data(), &done, out); });
            else simt_launch_fn(grid, 256, [&]() { k_icp_plane<true, false>(a.M, a.S, a.P, partials.data(), &done, out); });
        } else {
            if (search) simt_launch_fn(grid, 256, [&]() { k_icp_plane<false, true>(a.M, a.S, a.P, partials.data(), &done, out); });
            else simt_launch_fn(grid, 256, [&]() { k_icp_plane<false, false>(a.M, a.S, a.P, partials.data(), &done, out); });
        }
    };
    plane(true, out160);
    for (int i = 0; i < n; i++) {
        world[3 * (size_t)i] = w[i].x; world[3 * (size_t)i + 1] = w[i].y; world[3 * (size_t)i + 2] = w[i].z;
        int cnt = 0;
        for (int k = 0; k < 5; k++) {
            const int id = ids[(size_t)i * 5 + k];
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (id >= 0) {
                q = E->pool[id];
                cnt++;
            }
            near_xyz[15 * (size_t)i + 3 * k] = q.x; near_xyz[15 * (size_t)i + 3 * k + 1] = q.y; near_xyz[15 * (size_t)i + 3 * k + 2] = q.z;
        }
        near_cnt[i] = cnt;
    }
    if (pose2) {
        set_pose(pose2);
        plane(false, out160_2);
    }
    memcpy(selected, sel.data(), n);
    memcpy(normvec, nv.data(), (size_t)n * 16);
}

}  // extern "C"
