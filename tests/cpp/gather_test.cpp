// gather_test.cpp — sb_gather_results / sb_gather_candidates from a C++ host program, one process per GPU, no Python in
// the data path: usage  gather_test <rank> <world> <id_file>.  Rank 0 creates the NCCL unique id and writes it to
// id_file; the other ranks wait for it.  Every rank fabricates the results of the units it owns (unit u -> rank
// u % world), gathers, and checks every record of every unit; the same for per-rank candidate lists.
#include <nccl.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "slam_b200.h"

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static void fill(sb_icp_result& r, int u) {   // a result that depends on the unit only
    std::memset(&r, 0, sizeof(r));
    for (int i = 0; i < 16; ++i) r.transformation[i] = u * 100.0 + i;
    r.final_error = 0.001 * u;
    r.converged = u % 3 != 0;
    r.num_iterations = u % 17;
    r.history_len = r.num_iterations + 1;
    r.status = 0;
    for (int i = 0; i < r.history_len; ++i) r.error_history[i] = u + 0.5 * i;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int rank = std::atoi(argv[1]), world = std::atoi(argv[2]);
    const char* id_file = argv[3];
    int n_dev = 0;
    cudaGetDeviceCount(&n_dev);
    if (n_dev < 1) { std::printf("no GPU\n"); return 3; }
    const int dev = rank % n_dev;
    cudaSetDevice(dev);
    ncclUniqueId id;
    if (rank == 0) {
        ncclGetUniqueId(&id);
        std::string tmp = std::string(id_file) + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        std::fwrite(&id, sizeof(id), 1, f);
        std::fclose(f);
        std::rename(tmp.c_str(), id_file);
    } else {
        FILE* f = nullptr;
        for (int t = 0; t < 600 && !(f = std::fopen(id_file, "rb")); ++t) usleep(100000);
        if (!f) { std::printf("rank %d: no id file\n", rank); return 4; }
        if (std::fread(&id, sizeof(id), 1, f) != 1) return 4;
        std::fclose(f);
    }
    ncclComm_t comm;
    if (ncclCommInitRank(&comm, world, id, rank) != ncclSuccess) { std::printf("ncclCommInitRank failed\n"); return 5; }
    sb_ctx* ctx = nullptr;
    CHECK(sb_ctx_create(dev, nullptr, &ctx) == SB_OK);

    for (int n_total : {0, 1, 5, 64, 1001}) {   // ragged: 1001 % world != 0 for world 2, 4, 8
        std::vector<sb_icp_result> local, all((size_t)(n_total > 0 ? n_total : 1));
        for (int u = rank; u < n_total; u += world) { local.emplace_back(); fill(local.back(), u); }
        int s = sb_gather_results(ctx, comm, rank, world, local.empty() ? nullptr : local.data(), n_total, all.data());
        CHECK(s == SB_OK);
        for (int u = 0; u < n_total; ++u) {
            sb_icp_result want;
            fill(want, u);
            CHECK(std::memcmp(&want, &all[(size_t)u], sizeof(want)) == 0);
        }
    }
    {   // candidates: rank r offers (0.01 * (i * world + r) , entry i * world + r) plus one tie on the distance
        const int cap = 16, n_local = 10;
        std::vector<double> d((size_t)n_local), od((size_t)cap);
        std::vector<int32_t> e((size_t)n_local), oe((size_t)cap);
        for (int i = 0; i < n_local; ++i) { e[(size_t)i] = i * world + rank; d[(size_t)i] = i == 0 ? 0.0 : 0.01 * e[(size_t)i]; }
        int32_t cnt = -1;
        CHECK(sb_gather_candidates(ctx, comm, world, d.data(), e.data(), n_local, cap, od.data(), oe.data(), &cnt) == SB_OK);
        const int total = n_local * world;
        CHECK(cnt == (total < cap ? total : cap));
        // expected order: the `world` entries with distance 0 first (ascending entry: loop_closure.hpp:92 sorts pairs), then by distance
        for (int i = 0; i < cnt; ++i) {
            if (i < world) { CHECK(od[(size_t)i] == 0.0 && oe[(size_t)i] == i); }
            else { CHECK(oe[(size_t)i] == i && std::fabs(od[(size_t)i] - 0.01 * i) < 1e-15); }
        }
    }
    sb_ctx_destroy(ctx);
    ncclCommDestroy(comm);
    if (failures) { std::printf("gather_test rank %d: %d FAILURES\n", rank, failures); return 1; }
    std::printf("gather_test rank %d of %d: all checks passed\n", rank, world);
    return 0;
}
