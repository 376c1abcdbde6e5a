/*
 * yask_b200.h -- C ABI of the B200-native stencil engine (libyask_b200.so).
 *
 * This is the drop-in boundary for the hot path of intel/yask (yk_solution::run_solution and
 * the var/geometry calls around it).  The reference has no C plugin ABI: its boundary is the
 * C++ kernel API (yk_factory / yk_env / yk_solution / yk_var / yk_stats,
 * /root/reference/include/yask_kernel_api.hpp, aux/yk_solution_api.hpp, aux/yk_var_api.hpp),
 * implemented by a per-stencil libyask_kernel.<stencil>.<arch>.so.  Each entry point below is
 * what a binding of that C++ API needs from a device engine; the C++ mirror of the reference
 * API that sits on top of it lives in yask_b200/include/yask_kernel_api.hpp.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, a negative YB_E* code on
 *     failure, and yb_last_error() returns the message for the calling thread (the C++ layer
 *     turns it into yask::yask_exception, /root/reference/include/yask_common_api.hpp:125-179).
 *   - indices are GLOBAL ("overall problem") element indices exactly as in the reference's
 *     yk_var API (aux/yk_var_api.hpp:60-183): domain index 0 is the first point of the overall
 *     domain, halo points of the first rank are negative.
 *   - dims of a var are given in the var's declared order; the step dim (if any) comes first.
 *   - all data lives in device (HBM) memory; host buffers passed to slice calls are copied.
 *   - there is NO CPU fallback: without a CUDA device yb_solution_prepare() fails.
 */
#ifndef YASK_B200_H
#define YASK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YB_MAX_DIMS 5          /* step + up to 3 domain + misc */
#define YB_MAX_DOMAIN_DIMS 3
#define YB_NAME_LEN 64

enum {
    YB_OK = 0,
    YB_EINVAL = -1,    /* bad argument / unknown name  (reference: yask_exception) */
    YB_ESTATE = -2,    /* call not legal in this state, e.g. run before prepare (context.cpp:265-266) */
    YB_ERANGE = -3,    /* index outside the allocation / invalid step (yk_var.hpp:1084-1135) */
    YB_ECUDA = -4,     /* CUDA runtime/driver error, or no device */
    YB_ENOMEM = -5,
    YB_EUNSUPPORTED = -6
};

/* FP-contraction modes of the point-update arithmetic (see oracle/yask_oracle.c). */
enum {
    YB_FP_STRICT = 0,   /* IEEE mul/add in DSL order == reference built with -ffp-contract=off */
    YB_FP_FMA = 1,      /* canonical fused form */
    YB_FP_REF_GCC = 2   /* FMA pattern of the reference's default GCC build (default) */
};

typedef struct yb_solution yb_solution;   /* opaque; mirrors yk_solution (aux/yk_solution_api.hpp:82) */

/* Per-dim geometry of one var, all in elements.  Mirrors the getters of yk_var
 * (aux/yk_var_api.hpp:185-640) and SURVEY.md Appendix C. */
typedef struct yb_dim_info {
    char name[YB_NAME_LEN];
    int32_t kind;             /* 0 = step, 1 = domain, 2 = misc */
    int32_t domain_index;     /* for kind==1: position in the solution's domain dims, else -1 */
    int64_t rank_offset;      /* global index of this rank's first domain point */
    int64_t domain_size;      /* rank domain size (kind 1); alloc_t (kind 0); misc extent (kind 2) */
    int64_t left_halo, right_halo;
    int64_t left_pad, right_pad;   /* >= halo; alloc = left_pad + domain_size + right_pad */
    int64_t alloc_size;
    int64_t first_misc_index;
    int64_t stride;           /* element stride in device storage (0 for the step dim) */
} yb_dim_info;

typedef struct yb_var_info {
    char name[YB_NAME_LEN];
    int32_t num_dims;
    int32_t elem_bytes;
    int32_t has_step;
    int32_t step_alloc;           /* number of step slots (alloc_t) */
    int64_t first_valid_step, last_valid_step;
    int32_t is_output;            /* written by the solution */
    int32_t halo_exchange_l1_norm;
    int64_t slot_elems;           /* elements per step slot (incl. pads) */
    int64_t storage_bytes;
    yb_dim_info dims[YB_MAX_DIMS];
} yb_var_info;

/* Mirrors yk_stats (aux/yk_solution_api.hpp:1300-1348). */
typedef struct yb_stats {
    int64_t num_elements;         /* overall domain points */
    int64_t num_steps_done;
    int64_t num_writes_done;
    int64_t est_fp_ops_done;
    int64_t num_reads_done;
    double elapsed_secs;          /* device time inside yb_solution_run (CUDA events) */
    double halo_secs;             /* part of elapsed spent waiting on halo exchange */
    int64_t kernel_launches;      /* stencil kernels launched since last clear */
} yb_stats;

/* ---- library ------------------------------------------------------------------------------ */
const char* yb_version_string(void);              /* yk_factory::get_version_string, yask_kernel_api.hpp:90 */
const char* yb_last_error(void);
int yb_device_count(void);                        /* number of visible CUDA devices (0 if none) */
int yb_num_stencils(void);
const char* yb_stencil_name(int i);               /* registry of built-in solutions (REGISTER_SOLUTION names) */

/* ---- solution life cycle (yk_factory::new_solution, factory.cpp:51-105) --------------------- */
/* stencil: "iso3dfd" | ... ; radius <= 0 selects the stencil's default (iso3dfd: 8);
 * elem_bytes: 4 or 8 (0 = default 4). */
int yb_solution_create(yb_solution** out, const char* stencil, int radius, int elem_bytes);
int yb_solution_destroy(yb_solution* s);          /* yk_solution::end_solution + release */
const char* yb_solution_name(const yb_solution* s);           /* get_name */
const char* yb_solution_target(const yb_solution* s);         /* get_target: "sm_100a" */
int yb_solution_elem_bytes(const yb_solution* s);             /* get_element_bytes */
int yb_solution_num_domain_dims(const yb_solution* s);        /* get_num_domain_dims */
const char* yb_solution_domain_dim_name(const yb_solution* s, int i);
const char* yb_solution_step_dim_name(const yb_solution* s);

/* ---- settings, legal before prepare (aux/yk_solution_api.hpp:187-518) ------------------------ */
int yb_set_rank_domain_size(yb_solution* s, int dim, int64_t n);     /* set_rank_domain_size */
int yb_set_overall_domain_size(yb_solution* s, int dim, int64_t n);  /* set_overall_domain_size */
int yb_set_num_ranks(yb_solution* s, int dim, int64_t n);            /* set_num_ranks */
int yb_set_rank_index(yb_solution* s, int dim, int64_t i);           /* set_rank_index */
int yb_set_min_pad_size(yb_solution* s, int dim, int64_t n);         /* set_min_pad_size */
int64_t yb_get_rank_domain_size(const yb_solution* s, int dim);
int64_t yb_get_overall_domain_size(const yb_solution* s, int dim);
int64_t yb_get_num_ranks(const yb_solution* s, int dim);
int64_t yb_get_rank_index(const yb_solution* s, int dim);
int64_t yb_get_first_rank_domain_index(const yb_solution* s, int dim);  /* valid after prepare */
int64_t yb_get_last_rank_domain_index(const yb_solution* s, int dim);
/* Engine options ("-key value" strings of apply_command_line_options, soln_apis.cpp:285-313):
 *   fp_mode = 0|1|2, kernel = auto|tma|direct, lx = <planes per sweep chunk>, ...            */
int yb_set_option(yb_solution* s, const char* key, const char* value);
int yb_get_option(const yb_solution* s, const char* key, char* value, size_t value_len);
/* Run the kernels on this CUDA stream (a cudaStream_t cast to void*; NULL = the solution's own). */
int yb_set_stream(yb_solution* s, void* cuda_stream);

/* Host-only part of prepare (setup_rank, setup.cpp:462-503): derive this rank's domain size and
 * offset from the overall size / rank grid / rank index.  Needs no device; prepare() calls it too. */
int yb_solution_plan_geometry(yb_solution* s);

/* ---- prepare: rank geometry + device allocation (prepare_solution, soln_apis.cpp:137-249) ---- */
int yb_solution_prepare(yb_solution* s, int device);
int yb_solution_is_prepared(const yb_solution* s);

/* ---- vars (yk_solution::get_var/get_vars, yk_var getters) ----------------------------------- */
int yb_num_vars(const yb_solution* s);
int yb_var_index(const yb_solution* s, const char* name);     /* <0 if unknown */
int yb_var_info_get(const yb_solution* s, int var, yb_var_info* out);
/* User-created vars (yk_solution::new_var / new_fixed_size_var, aux/yk_solution_api.hpp:994-1089): not read
 * or written by the stencil kernels.  dim names matching the step dim / a domain dim take that role, any
 * other name is a misc dim.  sizes == NULL: sized like the solution (new_var); otherwise fixed sizes (one per
 * dim: steps / domain points / misc extent).  Returns the new var's index (>= 0) or a negative error. */
int yb_var_create(yb_solution* s, const char* name, int ndims, const char* const* dim_names, const int64_t* sizes);

/* yk_var::fuse_vars (aux/yk_var_api.hpp:1370-1397; /root/reference/src/kernel/lib/yk_var_apis.cpp:334-360): `var` of `s`
 * becomes another reference to the device storage of `src_var` of `src` (the same or another solution of this process);
 * layouts must be identical; storage `var` had is released when its last user lets go. */
int yb_var_fuse(yb_solution* s, int var, yb_solution* src, int src_var);
/* Set per-var geometry before prepare (set_halo_size / set_min_pad_size, yk_var_api.hpp:1180-1290). */
int yb_var_set_min_pad(yb_solution* s, int var, int dim, int64_t left, int64_t right);

/* Slice copies between a caller-owned HOST buffer (row-major in declared dim order, last dim
 * unit stride) and device storage: yk_var::set_elements_in_slice / get_elements_in_slice
 * (aux/yk_var_api.hpp:699-751, 876-961).  first/last have num_dims entries (step first).
 * Returns the number of elements copied in *n_done (may be NULL). */
int yb_var_set_slice(yb_solution* s, int var, const void* host_buf, const int64_t* first, const int64_t* last, int64_t* n_done);
int yb_var_get_slice(yb_solution* s, int var, void* host_buf, const int64_t* first, const int64_t* last, int64_t* n_done);
/* Same, but the buffer is DEVICE memory on the solution's device (no host round trip). */
int yb_var_set_slice_device(yb_solution* s, int var, const void* dev_buf, const int64_t* first, const int64_t* last, int64_t* n_done);
int yb_var_get_slice_device(yb_solution* s, int var, void* dev_buf, const int64_t* first, const int64_t* last, int64_t* n_done);
/* yk_solution::run_auto_tuner_now (aux/yk_solution_api.hpp:858-882): times the engine's launch variants over the rank
 * domain right now (no halo exchange; var contents are NOT preserved), keeps the fastest for later run_solution()
 * calls, clears the stats.  `report` (may be NULL) receives one line per trial. */
int yb_solution_auto_tune(yb_solution* s, char* report, size_t report_len);
/* In-run tuner (yk_solution::reset_auto_tuner, is_auto_tuner_enabled, option -auto_tune; aux/yk_solution_api.hpp:820-856):
 * while enabled, every run_solution() step executes with the next untried launch variant and is timed; when all variants
 * have their samples the fastest stays selected and the tuner switches itself off.  Variants compute identical bits, so the
 * steps taken while tuning are real steps. */
int yb_solution_reset_auto_tuner(yb_solution* s, int enable);
int yb_solution_is_auto_tuner_enabled(const yb_solution* s);
int yb_solution_auto_tuner_report(const yb_solution* s, char* report, size_t report_len);
/* Reductions over a slice, on the device, accumulated in double in a fixed order: yk_var::reduce_elements_in_slice
 * (aux/yk_var_api.hpp:984-1110; /root/reference/src/kernel/lib/yk_var.hpp:1367-1450).
 * out = {sum, sum of squares, product, max, min}; *n_done = number of elements reduced (may be NULL). */
int yb_var_reduce_slice(yb_solution* s, int var, const int64_t* first, const int64_t* last, double out[5], int64_t* n_done);
/* set_all_elements_same (writes every storage element incl. pads, yk_var.hpp:1786-1793). */
int yb_var_set_all_same(yb_solution* s, int var, double value);
/* set_elements_in_slice_same. */
int yb_var_set_slice_same(yb_solution* s, int var, double value, const int64_t* first, const int64_t* last, int64_t* n_done);
/* Deterministic synthetic data on the device: value = lo + (hi-lo)*u(hash(seed, salt, global idx))
 * over the rank's halo box of API step `step` (same function as yask_b200/synth.py). */
int yb_var_fill_hash(yb_solution* s, int var, int64_t step, uint32_t seed, uint32_t salt, double lo, double hi);
/* Same, with `shift[d]` (one entry per solution domain dim; NULL = zeros) added to the global index of every point:
 * a small stand-alone solution can then hold exactly the values a window of a larger (multi-rank) problem holds --
 * bench.py's halo check recomputes the rank interfaces this way. */
int yb_var_fill_hash_shifted(yb_solution* s, int var, int64_t step, uint32_t seed, uint32_t salt, double lo, double hi,
                             const int64_t* shift);
/* Order-independent 64-bit checksum (sum of per-element bit patterns mixed with the global index)
 * over the rank-domain box of API step `step`. */
int yb_var_checksum(yb_solution* s, int var, int64_t step, uint64_t* out);
/* Raw device pointer of a step slot (get_raw_storage_buffer, yk_var_api.hpp:1399-1437). */
int yb_var_device_ptr(yb_solution* s, int var, int64_t step, void** out);

/* Plain device->host copy of `bytes` bytes (for host snapshots of raw storage). */
int yb_copy_to_host(void* host_dst, const void* dev_src, size_t bytes);

/* ---- run (yk_solution::run_solution, context.cpp:220-624) ----------------------------------- */
int yb_solution_run(yb_solution* s, int64_t first_step, int64_t last_step);
int yb_solution_sync(yb_solution* s);              /* wait for all queued device work */
int yb_get_stats(yb_solution* s, yb_stats* out);   /* get_stats (syncs) */
int yb_clear_stats(yb_solution* s);                /* clear_stats */

/* ---- multi-GPU halo exchange, one process per GPU (replaces exchange_halos, halo.cpp:80-491) --- */
/* Each rank exports an opaque blob (CUDA IPC handles of its var storage + sync flags), the
 * launcher all-gathers the blobs (torch.distributed / MPI / files) and imports every
 * neighbour's blob.  After that the boundary kernels write halos straight into the peers' HBM
 * over NVLink; no further host communication is needed. */
int yb_halo_export_size(const yb_solution* s, size_t* nbytes);
int yb_halo_export(yb_solution* s, void* blob, size_t nbytes);
int yb_halo_import(yb_solution* s, int64_t peer_rank_linear, const void* blob, size_t nbytes);
int yb_halo_finalize(yb_solution* s);              /* after all imports: ready to run */
int yb_exchange_halos(yb_solution* s);             /* yk_solution::exchange_halos (public API) */

/* ---- job rendezvous without MPI or torch (replaces the MPI set-up of setup.cpp:169-524 and the harness's
 *      yk_env::global_barrier / sum_over_ranks, yask_kernel_api.hpp:238-293) ---------------------------------------
 * One process per GPU on one node; the ranks meet in a POSIX shared-memory mailbox named after `job_key` (NULL:
 * YASK_JOB_ID, else MASTER_ADDR:MASTER_PORT + the launcher's pid).  yb_comm_env_* read the rank / world size / local
 * rank a launcher exported (YASK_*, torchrun's RANK/WORLD_SIZE/LOCAL_RANK, Open MPI, PMI, Slurm). */
int yb_comm_env_rank(void);
int yb_comm_env_world(void);
int yb_comm_env_local_rank(void);
int yb_comm_init(int rank, int world, const char* job_key);
int yb_comm_rank(void);
int yb_comm_world(void);
int yb_comm_barrier(void);
int yb_comm_allgather(const void* mine, size_t nbytes, void* all /* world * nbytes */);
int yb_comm_sum_i64(int64_t v, int64_t* out);
int yb_comm_max_f64(double v, double* out);
int yb_comm_finalize(void);
/* export -> all-gather through the communicator -> import every neighbour -> finalize, for a prepared multi-rank solution
 * whose rank-grid position matches this process's job rank (row-major over the domain dims, x slowest). */
int yb_halo_connect(yb_solution* s);

#ifdef __cplusplus
}
#endif
#endif /* YASK_B200_H */
