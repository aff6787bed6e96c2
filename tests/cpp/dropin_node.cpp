// tests/cpp/dropin_node.cpp -- TEST: the node-side patch of INTEGRATION.md section 2 as a real translation unit.
//
// Compiles lidar_imu_init_b200/csrc/host/liinit_adapter.hpp + liinit_host.h against the reference's point type
// (pcl::PointXYZINormal, 48 bytes; here from oracle/shim, the stand-in the verbatim ikd-Tree is built with -- PCL is not installed)
// with Eigen::aligned_allocator as the node declares its PointVector (include/common_lib.h:37-41), and runs the per-scan
// sequence of laserMapping.cpp:921-1142 through it: Build -> UploadScan -> liinit_scan_update -> liinit_map_incremental ->
// Nearest_Search / size / validnum. tests/test_dropin_cpp.py compiles it on the CPU and, on a GPU, compares what it writes with
// the ctypes path on the same inputs.
//   usage: dropin_node <in.bin> <out.bin> [nranks]
//   nranks > 1: the SAME sequence as one process per GPU (fork), the multi-GPU block of INTEGRATION.md section 5: rank 0 draws the
//   communicator id (liinit_comm_unique_id), the others read it from a file (the application's "my_broadcast"), every rank calls
//   liinit_comm_init and then runs the unchanged per-scan code; rank r writes <out.bin>.r. No Python, no torch in those processes.
//   in : int32 n_map, n_scan, imu_en; float32 map[n_map*3], body[n_scan*3]; float64 state[612] (liinit_state), ds
//   out: float64 state[612]; int32 iterations, search_passes, effect_feat_num, n_add, n_nod, size, validnum, near_cnt;
//        float32 near_xyz[15], near_d2[5]
#include <pcl/point_types.h>
#include <Eigen/StdVector>

#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "liinit_host.h"   // pulls liinit_gpu.h
#include "liinit_adapter.hpp"

typedef pcl::PointXYZINormal PointType;                                        // include/common_lib.h:37
typedef std::vector<PointType, Eigen::aligned_allocator<PointType>> PointVector;   // include/common_lib.h:41
static_assert(sizeof(PointType) == 48, "PointXYZINormal layout");

static std::unique_ptr<liinit::DeviceMap<PointType, Eigen::aligned_allocator<PointType>>> gmap;   // replaces `KD_TREE ikdtree;` (:125)

static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

static int run_rank(const char* in_path, const std::string& out_path, int rank, int nranks, const std::string& id_path);

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const int nranks = argc > 3 ? atoi(argv[3]) : 1;
    if (nranks <= 1) return run_rank(argv[1], argv[2], 0, 1, "");
    const std::string id_path = std::string(argv[2]) + ".id";
    remove(id_path.c_str());
    std::vector<pid_t> kids;
    for (int r = 0; r < nranks; r++) {   // fork BEFORE any CUDA call: every child creates its own context on device r
        pid_t p = fork();
        if (p == 0) _exit(run_rank(argv[1], std::string(argv[2]) + "." + std::to_string(r), r, nranks, id_path));
        kids.push_back(p);
    }
    int bad = 0;
    for (pid_t p : kids) {
        int st = 0;
        waitpid(p, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
    }
    return bad;
}

static int run_rank(const char* in_path, const std::string& out_path, int rank, int nranks, const std::string& id_path) {
    FILE* f = fopen(in_path, "rb");
    if (!f) return 2;
    int hdr[3];
    if (!rd(f, hdr, sizeof(hdr))) return 2;
    const int n_map = hdr[0], n_scan = hdr[1], imu_en = hdr[2];
    std::vector<float> mxyz((size_t)n_map * 3), bxyz((size_t)n_scan * 3);
    liinit_state st;
    double ds = 0;
    if (!rd(f, mxyz.data(), mxyz.size() * 4) || !rd(f, bxyz.data(), bxyz.size() * 4) || !rd(f, &st, sizeof(st)) || !rd(f, &ds, 8)) return 2;
    (void)nranks;
    fclose(f);
    static_assert(sizeof(liinit_state) == 612 * sizeof(double), "liinit_state is 612 doubles");

    PointVector feats_down_world(n_map), feats_down_body(n_scan);
    for (int i = 0; i < n_map; i++) { feats_down_world[i].x = mxyz[3 * i]; feats_down_world[i].y = mxyz[3 * i + 1]; feats_down_world[i].z = mxyz[3 * i + 2]; }
    for (int i = 0; i < n_scan; i++) {
        feats_down_body[i].x = bxyz[3 * i]; feats_down_body[i].y = bxyz[3 * i + 1]; feats_down_body[i].z = bxyz[3 * i + 2];
        feats_down_body[i].intensity = (float)i;   // the other 9 floats of the struct must not matter
        feats_down_body[i].curvature = 0.5f * i;
    }
    try {
        liinit_config cfg{};
        cfg.filter_size_map = (float)ds;
        cfg.max_map_points = n_map * 2 + 100000;
        cfg.max_scan_points = n_scan + 16;
        cfg.device_id = rank;
        gmap.reset(new liinit::DeviceMap<PointType, Eigen::aligned_allocator<PointType>>(cfg));
        if (nranks > 1) {
            unsigned char id[LIINIT_COMM_ID_BYTES];
            if (rank == 0) {
                if (liinit_comm_unique_id(id) != LIINIT_OK) { fprintf(stderr, "liinit_comm_unique_id failed\n"); return 1; }
                const std::string tmp = id_path + ".tmp";
                FILE* o = fopen(tmp.c_str(), "wb");
                fwrite(id, 1, sizeof(id), o);
                fclose(o);
                rename(tmp.c_str(), id_path.c_str());      // "my_broadcast": a file that appears atomically
            } else {
                FILE* i = nullptr;
                for (int tries = 0; tries < 600 && !(i = fopen(id_path.c_str(), "rb")); tries++) usleep(100000);
                if (!i || fread(id, 1, sizeof(id), i) != sizeof(id)) { fprintf(stderr, "rank %d: no communicator id\n", rank); return 1; }
                fclose(i);
            }
            if (liinit_comm_init(gmap->ctx(), id, nranks, rank) != LIINIT_OK) {
                fprintf(stderr, "liinit_comm_init: %s\n", liinit_last_error(gmap->ctx()));
                return 1;
            }
        }
        gmap->Build(feats_down_world);                       // ikdtree.Build(feats_down_world->points) (:928)
        gmap->UploadScan(feats_down_body);                   // feats_down_body, once per scan
        liinit_scan_stats ss;
        if (liinit_scan_update(gmap->ctx(), &st, 5, imu_en, &ss) != LIINIT_OK) {
            fprintf(stderr, "liinit: %s\n", liinit_last_error(gmap->ctx()));
            return 1;
        }
        int n_add = 0, n_nod = 0;
        if (liinit_map_incremental(gmap->ctx(), st.rot_end, st.pos_end, st.offset_R_L_I, st.offset_T_L_I, ds, 1, &n_add, &n_nod) != LIINIT_OK) {
            fprintf(stderr, "liinit: %s\n", liinit_last_error(gmap->ctx()));
            return 1;
        }
        PointVector near;
        std::vector<float> d2;
        PointType q = feats_down_world[n_map / 2];
        q.x += 0.05f;
        gmap->Nearest_Search(q, 5, near, d2);
        int tail[8] = {ss.iterations, ss.search_passes, ss.effect_feat_num, n_add, n_nod, gmap->size(), gmap->validnum(), (int)near.size()};
        float nx[15] = {0}, nd[5] = {0};
        for (size_t j = 0; j < near.size() && j < 5; j++) { nx[3 * j] = near[j].x; nx[3 * j + 1] = near[j].y; nx[3 * j + 2] = near[j].z; nd[j] = d2[j]; }
        FILE* o = fopen(out_path.c_str(), "wb");
        if (!o) return 2;
        fwrite(&st, sizeof(st), 1, o);
        fwrite(tail, sizeof(tail), 1, o);
        fwrite(nx, sizeof(nx), 1, o);
        fwrite(nd, sizeof(nd), 1, o);
        fclose(o);
        gmap.reset();
    } catch (const std::exception& e) {
        fprintf(stderr, "dropin_node: %s\n", e.what());
        return 1;
    }
    return 0;
}
