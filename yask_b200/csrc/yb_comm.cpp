// Rendezvous of the ranks of one job on ONE node without MPI or torch: a POSIX shared-memory mailbox.
//
// The reference sets its ranks up with MPI (/root/reference/src/kernel/lib/setup.cpp:169-524: MPI_Comm_rank/size, the
// neighbour table, MPI-3 shared-memory windows) and its harness relies on MPI barriers and reductions
// (/root/reference/src/kernel/yask_main.cpp:440-478, yk_env::global_barrier / sum_over_ranks).  The B200 engine needs the
// host side only ONCE per solution -- to hand every rank its neighbours' CUDA-IPC handles (yb_halo.cu) -- plus the
// harness's barriers and integer reductions; the data plane is peer stores over NVLink.  So the whole "communicator" is a
// file under /dev/shm: a sense-reversing barrier (two atomics) and one slot per rank for all-gathers.  One process per
// GPU, started by any launcher that gives each process its rank and the world size (RANK / WORLD_SIZE as torchrun sets them,
// or OMPI_COMM_WORLD_*, PMI_*, SLURM_*).
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <string>

#include "yb_core.h"

namespace yb {

namespace {

constexpr uint64_t COMM_MAGIC = 0x59423230434f4d4dull;   // "YB20COMM"
constexpr size_t SLOT_BYTES = 64 * 1024;

struct CommHeader {
    std::atomic<uint64_t> magic;       // set last by rank 0
    uint64_t world;
    std::atomic<uint64_t> arrived;     // barrier: ranks that reached the current generation
    std::atomic<uint64_t> generation;  // barrier: completed generations
    std::atomic<uint64_t> attached;    // ranks that mapped the file (rank 0 unlinks the name once all have)
    char pad[64];
};

struct Comm {
    int rank = 0, world = 1;
    CommHeader* hdr = nullptr;
    char* slots = nullptr;
    size_t map_bytes = 0;
    std::string name;
};
Comm g_comm;

void nap() {
    struct timespec ts = {0, 50000};   // 50 us
    nanosleep(&ts, nullptr);
}

int env_int(const char* const* names, int dflt) {
    for (; *names; names++)
        if (const char* v = getenv(*names)) return atoi(v);
    return dflt;
}

}  // namespace

}  // namespace yb

using namespace yb;

extern "C" {

int yb_comm_env_rank(void) {
    static const char* n[] = {"YASK_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", nullptr};
    return env_int(n, 0);
}
int yb_comm_env_world(void) {
    static const char* n[] = {"YASK_WORLD_SIZE", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", nullptr};
    return std::max(1, env_int(n, 1));
}
int yb_comm_env_local_rank(void) {
    static const char* n[] = {"YASK_LOCAL_RANK", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", nullptr};
    return env_int(n, yb_comm_env_rank());
}

int yb_comm_rank(void) { return g_comm.rank; }
int yb_comm_world(void) { return g_comm.world; }

// job_key: any string all ranks of the job agree on and other jobs on the node do not use; NULL derives it from
// YASK_JOB_ID, else MASTER_ADDR:MASTER_PORT plus the launcher's pid (the ranks' common parent process).
int yb_comm_init(int rank, int world, const char* job_key) {
    if (g_comm.hdr) return (rank == g_comm.rank && world == g_comm.world) ? 0 : set_error(YB_ESTATE, "communicator already initialised with another rank/world");
    if (world < 1 || rank < 0 || rank >= world) return set_error(YB_EINVAL, "bad rank %d / world %d", rank, world);
    g_comm.rank = rank; g_comm.world = world;
    if (world == 1) return 0;
    std::string key;
    if (job_key && *job_key) key = job_key;
    else if (const char* j = getenv("YASK_JOB_ID")) key = j;
    else {
        const char* a = getenv("MASTER_ADDR");
        const char* p = getenv("MASTER_PORT");
        key = std::string(a ? a : "local") + "_" + (p ? p : "0") + "_" + std::to_string((long long)getppid());
    }
    for (auto& c : key)
        if (!((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_' || c == '-')) c = '_';
    g_comm.name = "/yask_b200_" + key;
    g_comm.map_bytes = sizeof(CommHeader) + size_t(world) * SLOT_BYTES;
    int fd = -1;
    if (rank == 0) {
        shm_unlink(g_comm.name.c_str());      // a stale mailbox of a crashed job with the same key
        fd = shm_open(g_comm.name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0) return set_error(YB_EINVAL, "shm_open(%s) failed: %s", g_comm.name.c_str(), strerror(errno));
        if (ftruncate(fd, off_t(g_comm.map_bytes)) != 0) { close(fd); return set_error(YB_ENOMEM, "ftruncate(%s) failed: %s", g_comm.name.c_str(), strerror(errno)); }
    } else {
        for (int tries = 0; fd < 0; tries++) {      // wait for rank 0 (up to ~120 s)
            fd = shm_open(g_comm.name.c_str(), O_RDWR, 0600);
            if (fd >= 0) {
                struct stat sb;
                if (fstat(fd, &sb) == 0 && size_t(sb.st_size) >= g_comm.map_bytes) break;
                close(fd); fd = -1;
            }
            if (tries > 2400000) return set_error(YB_ESTATE, "rank %d: mailbox %s never appeared (is rank 0 running with the same MASTER_PORT / YASK_JOB_ID?)", rank, g_comm.name.c_str());
            nap();
        }
    }
    void* m = mmap(nullptr, g_comm.map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return set_error(YB_ENOMEM, "mmap of %s failed: %s", g_comm.name.c_str(), strerror(errno));
    g_comm.hdr = static_cast<CommHeader*>(m);
    g_comm.slots = static_cast<char*>(m) + sizeof(CommHeader);
    if (rank == 0) {
        g_comm.hdr->world = uint64_t(world);
        g_comm.hdr->arrived.store(0); g_comm.hdr->generation.store(0); g_comm.hdr->attached.store(0);
        g_comm.hdr->magic.store(COMM_MAGIC, std::memory_order_release);
    } else {
        for (long tries = 0; g_comm.hdr->magic.load(std::memory_order_acquire) != COMM_MAGIC; tries++) {
            if (tries > 2400000) return set_error(YB_ESTATE, "rank %d: mailbox %s was never initialised by rank 0", rank, g_comm.name.c_str());
            nap();
        }
        if (g_comm.hdr->world != uint64_t(world)) return set_error(YB_EINVAL, "rank %d: mailbox %s belongs to a job of %llu ranks, not %d", rank, g_comm.name.c_str(), (unsigned long long)g_comm.hdr->world, world);
    }
    // once everybody has the mapping the name is no longer needed (nothing is left behind if the job crashes later)
    if (g_comm.hdr->attached.fetch_add(1) + 1 == uint64_t(world)) shm_unlink(g_comm.name.c_str());
    return 0;
}

int yb_comm_barrier(void) {
    if (g_comm.world == 1) return 0;
    if (!g_comm.hdr) return set_error(YB_ESTATE, "communicator not initialised");
    CommHeader* h = g_comm.hdr;
    const uint64_t gen = h->generation.load(std::memory_order_acquire);
    if (h->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == uint64_t(g_comm.world)) {
        h->arrived.store(0, std::memory_order_relaxed);
        h->generation.store(gen + 1, std::memory_order_release);
    } else {
        for (long tries = 0; h->generation.load(std::memory_order_acquire) == gen; tries++) {
            if (tries > 12000000) return set_error(YB_ESTATE, "rank %d: barrier timed out (a rank of the job died?)", g_comm.rank);
            if (tries > 2000) nap();
        }
    }
    return 0;
}

// all[r * nbytes .. ) = rank r's `mine`
int yb_comm_allgather(const void* mine, size_t nbytes, void* all) {
    if (!mine || !all) return set_error(YB_EINVAL, "null argument");
    if (g_comm.world == 1) { memcpy(all, mine, nbytes); return 0; }
    if (!g_comm.hdr) return set_error(YB_ESTATE, "communicator not initialised");
    for (size_t off = 0; off < nbytes || off == 0; off += SLOT_BYTES) {
        const size_t n = std::min(SLOT_BYTES, nbytes - off);
        memcpy(g_comm.slots + size_t(g_comm.rank) * SLOT_BYTES, static_cast<const char*>(mine) + off, n);
        if (int rc = yb_comm_barrier()) return rc;
        for (int r = 0; r < g_comm.world; r++) memcpy(static_cast<char*>(all) + size_t(r) * nbytes + off, g_comm.slots + size_t(r) * SLOT_BYTES, n);
        if (int rc = yb_comm_barrier()) return rc;      // slots may be overwritten again
        if (nbytes == 0) break;
    }
    return 0;
}

int yb_comm_sum_i64(int64_t v, int64_t* out) {
    if (!out) return set_error(YB_EINVAL, "null argument");
    if (g_comm.world == 1) { *out = v; return 0; }
    std::string buf(size_t(g_comm.world) * sizeof(int64_t), '\0');
    if (int rc = yb_comm_allgather(&v, sizeof v, &buf[0])) return rc;
    int64_t s = 0;
    for (int r = 0; r < g_comm.world; r++) { int64_t x; memcpy(&x, &buf[size_t(r) * sizeof x], sizeof x); s += x; }
    *out = s;
    return 0;
}

int yb_comm_max_f64(double v, double* out) {
    if (!out) return set_error(YB_EINVAL, "null argument");
    if (g_comm.world == 1) { *out = v; return 0; }
    std::string buf(size_t(g_comm.world) * sizeof(double), '\0');
    if (int rc = yb_comm_allgather(&v, sizeof v, &buf[0])) return rc;
    double m = v;
    for (int r = 0; r < g_comm.world; r++) { double x; memcpy(&x, &buf[size_t(r) * sizeof x], sizeof x); m = std::max(m, x); }
    *out = m;
    return 0;
}

int yb_comm_finalize(void) {
    if (g_comm.hdr) {
        munmap(g_comm.hdr, g_comm.map_bytes);
        g_comm.hdr = nullptr; g_comm.slots = nullptr;
    }
    g_comm.rank = 0; g_comm.world = 1;
    return 0;
}

// Wire a prepared multi-rank solution to its neighbours through the communicator: export this rank's blob
// (CUDA-IPC handles of the var storage and flag words), all-gather, import every other rank's blob, finalize.
// Linear rank of a rank-grid position = row-major over the domain dims (x slowest), as yb_halo.cu numbers its peers.
int yb_halo_connect(yb_solution* s) {
    if (!s) return set_error(YB_EINVAL, "null solution");
    int64_t nr = 1, lin = 0;
    const int ndd = yb_solution_num_domain_dims(s);
    for (int d = 0; d < ndd; d++) { nr *= yb_get_num_ranks(s, d); lin = lin * yb_get_num_ranks(s, d) + yb_get_rank_index(s, d); }
    if (nr <= 1) return 0;
    if (nr != g_comm.world) return set_error(YB_EINVAL, "the solution's rank grid has %lld ranks but the job has %d", (long long)nr, g_comm.world);
    if (lin != g_comm.rank) return set_error(YB_EINVAL, "rank-grid position of this process is linear rank %lld, its job rank is %d", (long long)lin, g_comm.rank);
    size_t nb = 0;
    if (int rc = yb_halo_export_size(s, &nb)) return rc;
    std::string mine(nb, '\0'), all(nb * size_t(g_comm.world), '\0');
    if (int rc = yb_halo_export(s, &mine[0], nb)) return rc;
    if (int rc = yb_comm_allgather(mine.data(), nb, &all[0])) return rc;
    for (int r = 0; r < g_comm.world; r++) {
        if (r == g_comm.rank) continue;
        if (int rc = yb_halo_import(s, r, all.data() + size_t(r) * nb, nb)) return rc;
    }
    if (int rc = yb_halo_finalize(s)) return rc;
    return yb_comm_barrier();     // nobody starts storing into a peer before every peer has mapped everything
}

}  // extern "C"
