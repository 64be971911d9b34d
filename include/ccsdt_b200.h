/*
 * ccsdt_b200.h -- C ABI of the B200-native fused CCSD(T) triples driver.
 *
 * Drop-in boundary for ExaChem's (T) hot path.  Each entry point names the reference interface it
 * replaces (paths relative to the ExaChem source tree):
 *
 *   ccsdt_set_space      <- MO("occ"/"virt"/"occ_alpha"/"virt_alpha").num_tiles(), MO.input_tile_sizes(),
 *                           k_spin, k_evl_sorted, is_restricted arguments of
 *                           CCSD_T_Fused_Driver<T>::execute  exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp:73-81,112-126
 *   ccsdt_tiles          <- Cholesky_2E_Util::setupMOIS(ec, chem_env, triples=true)
 *                           exachem/cholesky/cholesky_2e.cpp:186-190,230-279 and k_spin of
 *                           exachem/cc/ccsd_t/ccsd_t.cpp:245-249
 *   ccsdt_enumerate      <- the task loops of execute        ccsd_t_fused_driver.hpp:368-395,478
 *   ccsdt_task_terms     <- ccsd_t_data_{s1,d1,d2}_info_only ccsd_t_all_fused_{singles,doubles1,doubles2}.hpp
 *   ccsdt_count_ops      <- calculate_performance_ops        ccsd_t_fused_driver.hpp:83-87,548-638
 *                           (+ helper_calculate_num_ops      fused_common.hpp:131-261)
 *   ccsdt_put_dense / ccsdt_put_block / ccsdt_set_fetch
 *                        <- Tensor<T>::get on d_t1, d_t2, d_v2.{v2ijab,v2ijka,v2iabc}
 *                           ccsd_t_all_fused_singles.hpp:200,304; ..._doubles1.hpp:222,237,282;
 *                           ..._doubles2.hpp:215,230,335  (and the six LRUCache arguments: the HBM block
 *                           store replaces them)
 *   ccsdt_put_cholesky   <- setupV2Tensors (the caller's step before execute)  exachem/cholesky/v2tensors.cpp:52-90,
 *                           exachem/cc/ccsd_t/ccsd_t.cpp:168-193
 *   ccsdt_run / ccsdt_run_tasks <- execute's task loop + ccsd_t_fully_fused_none_df_none_task
 *                           ccsd_t_all_fused.hpp:77-286 + the kernel launcher ccsd_t_all_fused_gpu.cu:2571
 *                           + hostEnergyReduce ccsd_t_all_fused.hpp:19-32; returns the rank-partial
 *                           (energy1 = E[T], energy2 = E(T)) the caller reduces at ccsd_t.cpp:262-263
 *   ccsdt_set_task_counter <- AtomicCounterGA (allocate / fetch_add / deallocate) ccsd_t_fused_driver.hpp:169-172,456,541
 *   ccsdt_check_memory   <- check_memory_req                 exachem/cc/ccsd_t/hybrid.cpp:19-41
 *   ccsdt_estimate_memory <- the memory summary printed before the (T) loop   exachem/cc/ccsd_t/ccsd_t.cpp:95-152
 *   options.exec_tilesize, ccsdt_exec_tiles, ccsdt_make_exec_tiles
 *                        <- the caller's re-tiling of T1/T2/V2 for ccsdt_tilesize   exachem/cc/ccsd_t/ccsd_t.cpp:168-193
 *                           (execution tiles are cut inside the library; blocks are still requested in the caller's tiling)
 *   ccsdt_task_counter_open / _close <- AtomicCounterGA allocate / deallocate among the ranks of a node
 *                           ccsd_t_fused_driver.hpp:169-172,541
 *   ccsdt_share_attach / _detach <- every rank's own Tensor<T>::get of the same block and its per-rank host caches
 *                           exachem/cc/ccsd_t/ccsd_t.cpp:236-241 (one fetch per node, GPU-to-GPU copies for the other ranks)
 *   ccsdt_comm_unique_id / _init / _allreduce / _destroy <- ec.pg().reduce(&energy1 / &energy2, ReduceOp::sum, 0)
 *                           exachem/cc/ccsd_t/ccsd_t.cpp:262-263 (ncclAllReduce of the two energies)
 *   ccsdt_clear_blocks, ccsdt_release_cached <- the per-call allocation and release of all staging memory in execute
 *                           ccsd_t_fused_driver.hpp:197-250,495-530
 *
 * Conventions: plain pointers and sizes, no C++ or torch types.  Every function returns 0 on success
 * and a non-zero code on failure; ccsdt_last_error() gives the message.  (The reference's convention
 * is fatal: CUDA_SAFE -> exit(100), ccsd_t_common.hpp:26-31; the C++ adapter in
 * include/ccsd_t_fused_driver_b200.hpp turns non-zero into tamm_terminate.)  There is no CPU
 * fallback: ccsdt_create fails when no CUDA device is usable.
 *
 * A context is single-caller (one host thread), bound to one GPU.  One process per GPU; rank /
 * nranks in ccsdt_options select this process's share of the task list.
 */
#ifndef CCSDT_B200_H
#define CCSDT_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CCSDT_API __attribute__((visibility("default")))
#else
#define CCSDT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ccsdt_ctx ccsdt_ctx;

/* tensor ids; block ids are tile indices inside the tensor's own sub-space, exactly what the
 * reference passes to Tensor::get (a virtual tile p is passed as p - noab) */
enum {
  CCSDT_T1     = 0, /* d_t1   [V][O]          */
  CCSDT_T2     = 1, /* d_t2   [V][V][O][O]    */
  CCSDT_V_IJAB = 2, /* v2ijab [O][O][V][V]    */
  CCSDT_V_IJKA = 3, /* v2ijka [O][O][O][V]    */
  CCSDT_V_IABC = 4  /* v2iabc [O][V][V][V]    */
};

/* DMMA: the product kernel (TMA producer warp + FP64 DMMA consumer warps); SIMPLE: diagnostic
 * one-thread-per-element FMA kernel over the same staged panels */
enum { CCSDT_KERNEL_DMMA = 0, CCSDT_KERNEL_SIMPLE = 1 };

typedef struct ccsdt_options {
  int32_t kernel;        /* CCSDT_KERNEL_DMMA (product) or CCSDT_KERNEL_SIMPLE (diagnostic FMA kernel) */
  int32_t sub[3];        /* CTA box = (2*sub[0], 2*sub[1], 2*sub[2], 8, 8, 8) over (h1,h2,h3,p4,p5,p6);
                            each entry 1..3, product <= 3 (4, 8 or 12 DMMA warps + 1 TMA warp).  0 = default (1,1,1) */
  int32_t stages;        /* TMA ring depth, 0 = as many as fit */
  int32_t ctas_per_sm;   /* 0 = default for the box */
  int32_t rank, nranks;  /* this process's share of the task list (static cost-balanced split) */
  int32_t overlap;       /* 1 = stage task n+1 while task n computes (default), 0 = serial */
  int32_t stagger;       /* 1 = de-phase the CTAs sharing an SM by a fraction of a box time (default), 0 = off */
  int32_t verbose;
  int32_t symmetry;      /* 1 (default) = when two hole (particle) tiles of a task coincide, evaluate only the CTA
                            boxes with ascending box coordinates along those indices and weight them by their
                            permutation multiplicity (t3 is antisymmetric in same-spin indices, so d*d/D and
                            d*(d+s)/D are symmetric; the reference evaluates every element and lets `factor`
                            undo the over-count, ccsd_t_fused_driver.hpp:387-395); 0 = evaluate every box */
  int32_t exec_tilesize; /* execution tiling (the tiles the task list, the panels and the kernel work on):
                            0 (default) = the caller's tiles, i.e. the reference's task list and per-task energies;
                            > 0 = re-cut every spin block into near-equal tiles of about this extent, multiples of the CTA
                            box (8 particles / 2 holes), so that ragged tiles (ts 28: 28 -> 32 per particle index, x1.49
                            executed DMMAs) and tiny tiles (ts 16) cost nothing; blocks are still fetched in the caller's
                            tiling and cut or merged by the panel build; -1 = auto (re-cut only when the caller's particle
                            tiles are not multiples of 8 or are smaller than 24).  Task ids of ccsdt_run / ccsdt_run_tasks
                            then index the execution task list (ccsdt_exec_tiles + ccsdt_enumerate); totals are unchanged */
  int32_t prefetch_tasks; /* static hand-out + fetch callback: blocks of up to this many upcoming tasks are fetched while
                            the host would otherwise wait for the GPU; 0 = default (as far ahead as the block budget
                            allows), -1 = off */
  int64_t block_budget_bytes; /* HBM the block store may hold before it evicts least-recently-used blocks;
                            0 = what is free after the panel pools are allocated, minus a reserve */
  int32_t watchdog_ms;   /* a pipeline wait inside the fused kernel that lasts longer than this traps (reported as a CUDA
                            error) instead of hanging the GPU; 0 = default (20 000 ms of %globaltimer), -1 = never */
  int32_t check_symmetry; /* with symmetry = 1: 0 (default) = verify on the device, once per block / dense tensor as it arrives,
                            that T2, v2ijab (both index pairs), v2ijka (i,j) and v2iabc (b,c) are antisymmetric where that can be
                            seen inside one block (blocks whose two tiles coincide; whole tensors for ccsdt_put_dense) -- the run
                            fails with a message instead of returning an energy built on a wrong assumption, and keeps failing
                            until ccsdt_set_space replaces the operands; -1 = skip the check */
} ccsdt_options;

typedef struct ccsdt_stats {
  int64_t tasks_run;          /* kernel tasks executed by this rank */
  int64_t kernel_launches;    /* launches of this library's kernels */
  double  counted_flops;      /* total_num_ops share of the tasks run (reference's count) */
  double  seconds_total;      /* wall time of ccsdt_run */
  double  seconds_kernel;     /* CUDA-event time of the fused kernel launches (sum) */
  double  seconds_staging;    /* CUDA-event time of the panel-build launches (sum) */
  int64_t h2d_bytes, d2h_bytes; /* h2d includes the uploads (ccsdt_put_*) made since the previous run */
  int64_t blocks_fetched;     /* fetch-callback invocations */
  double  evaluated_flops;    /* counted_flops x (CTA boxes evaluated / CTA boxes of the tile): what the fused kernel
                                 had to execute after the symmetry reduction (== counted_flops with symmetry = 0) */
  int64_t blocks_evicted;     /* blocks the LRU dropped from the HBM block store during the run */
  double  seconds_fetch;      /* host time spent inside the fetch callback (the caller's Tensor::get) */
  double  seconds_host_wait;  /* host time blocked on the GPU (buffer reuse, ring space, final synchronisation) */
  double  executed_flops;     /* flops of the DMMAs the fused kernel issued for the evaluated boxes: evaluated_flops
                                 plus the padding of ragged tiles up to the CTA box and of K up to 4 */
  int64_t blocks_from_peers;  /* blocks copied from another rank's HBM through the node-shared directory (ccsdt_share_attach) */
  int64_t peer_bytes;         /* ... and their bytes (device to device over NVLink; not counted in h2d_bytes) */
} ccsdt_stats;

/* delivers one UNSORTED row-major block, i.e. what Tensor<T>::get(bid, buf) returns */
typedef int (*ccsdt_fetch_fn)(void* user, int tensor, const uint32_t bid[4], double* dst, size_t n);

/* ccsdt_destroy parks the context's device resources (streams, panel pools, the pinned fetch ring, the private memory
 * pool) for the next ccsdt_create on the same device of this process -- CCSD_T_Fused_Driver::execute creates and
 * destroys a context per call, and on a 0.2 s job allocation would otherwise cost as much as the work.
 * ccsdt_release_cached frees what is parked; CCSDT_B200_CACHE=0 in the environment disables parking. */
CCSDT_API int         ccsdt_create(ccsdt_ctx** out, int device); /* device < 0: the calling thread's current CUDA device */
CCSDT_API int         ccsdt_destroy(ccsdt_ctx* ctx);
CCSDT_API int         ccsdt_release_cached(void);
CCSDT_API const char* ccsdt_last_error(const ccsdt_ctx* ctx); /* ctx may be NULL: error of the last failed create */
CCSDT_API int         ccsdt_default_options(ccsdt_options* opt);
CCSDT_API int         ccsdt_set_options(ccsdt_ctx* ctx, const ccsdt_options* opt);

/* host-only helpers (no GPU needed; ctx-free) */
CCSDT_API int     ccsdt_tiles(int64_t n_occ_alpha, int64_t n_occ_beta, int64_t n_vir_alpha, int64_t n_vir_beta,
                    int64_t tilesize, int64_t* k_range, int32_t* k_spin, int32_t counts[4], int cap);
CCSDT_API int64_t ccsdt_enumerate(int noab, int nvab, const int32_t* k_spin, int is_restricted, int64_t* tasks7,
                        double* factors, int64_t cap, int64_t* n_outer);
CCSDT_API int     ccsdt_task_terms(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range,
                         int is_restricted, const int64_t task[6], uint8_t* s1_on /*9*/,
                         uint8_t* d1_on /*9*noab*/, uint8_t* d2_on /*9*nvab*/);
CCSDT_API int     ccsdt_count_ops(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range,
                        int is_restricted, long double* total_num_ops);
CCSDT_API int     ccsdt_partition(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range,
                        int is_restricted, int nranks, int32_t* owner /* one per kernel task */,
                        int64_t cap);
CCSDT_API int     ccsdt_check_memory(int tilesize, int nbf, size_t gpu_bytes, size_t* required);
/* host-only: the HBM the path will use for a tile space -- the counterpart of the memory summary the reference prints before
 * the (T) loop for its host-side staging buffers and block caches (exachem/cc/ccsd_t/ccsd_t.cpp:95-152).  `target` is
 * options.exec_tilesize (0: the caller's tiles).  All sizes in bytes. */
typedef struct ccsdt_memory_estimate {
  int64_t exec_max_hole_tile, exec_max_particle_tile; /* largest execution tiles */
  int64_t panel_bytes;      /* K-major operand panels of the two staging buffers */
  int64_t s1_bytes;         /* staged s1 operands of the two staging buffers */
  int64_t task_block_bytes; /* upper bound of the tensor blocks one task reads */
  int64_t tensor_bytes[5];  /* spin-conserving blocks of T1, T2, v2ijab, v2ijka, v2iabc (a fully resident block store) */
  int64_t minimum_bytes;    /* panels + s1 + the blocks of two tasks: below this the LRU thrashes */
} ccsdt_memory_estimate;
CCSDT_API int     ccsdt_estimate_memory(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                                        int target, ccsdt_memory_estimate* out);
/* permutation weight the fused kernel gives CTA box `box` (box coordinates along h1,h2,h3,p4,p5,p6) under
 * symmetry bits `sym` (bit 0: tiles h1b==h2b, 1: h2b==h3b, 2: p4b==p5b, 3: p5b==p6b); 0 = the box is the
 * mirror image of an evaluated one and is skipped.  Summed over a tile's boxes the weights count every box once.
 * Bit 4 (hole boxes are 2 wide, set together with bits 0 and 1) also skips the boxes on the triple hole diagonal:
 * all their elements repeat a hole index, where t3 vanishes by antisymmetry. */
CCSDT_API int     ccsdt_box_weight(int sym, const int32_t box[6]);

/* problem definition */
CCSDT_API int ccsdt_set_space(ccsdt_ctx* ctx, int noa, int nob, int nva, int nvb, const int64_t* k_range,
                    const int32_t* k_spin, const double* evl, int is_restricted);
/* the execution tiling in force (== the tiles of ccsdt_set_space unless options.exec_tilesize re-cut them); returns the
 * number of tiles, or minus that number when cap is too small.  counts = tiles per (occ a, occ b, virt a, virt b) */
CCSDT_API int     ccsdt_exec_tiles(const ccsdt_ctx* ctx, int64_t* k_range, int32_t* k_spin, int32_t counts[4], int cap);
CCSDT_API int64_t ccsdt_num_tasks(const ccsdt_ctx* ctx); /* kernel tasks of the execution task list */
/* host-only: the execution tiling ccsdt_set_space + options.exec_tilesize = target would produce */
CCSDT_API int     ccsdt_make_exec_tiles(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                                        int target, int64_t* out_range, int32_t* out_spin, int32_t counts[4], int cap);
/* host-only (CPU tests of the re-tiling logic): the canonical storage blocks one execution-tile block request
 * exec_bid of `tensor` is cut into.  22 int64 per piece: bid[4] of the canonical storage block, sign (+1/-1 of the
 * canonicalisation), perm[4] (requested dim d is dim perm[d] of the canonical block), store_off[4], exec_off[4], len[4]
 * (per requested dim), elems of the storage block.  Returns the number of pieces. */
CCSDT_API int64_t ccsdt_split_request(int noa, int nob, int nva, int nvb, const int64_t* store_range, const int32_t* store_spin,
                                      const int64_t* exec_range, const int32_t* exec_counts, int tensor,
                                      const uint32_t exec_bid[4], int64_t* pieces /* 22 per piece */, int64_t cap);

/* operand supply (choose one per tensor).  ccsdt_put_dense copies only the spin-conserving blocks of the dense host array
 * (T1: s_a = s_i; four-index tensors: s_0 + s_1 = s_2 + s_3 -- the only blocks any task reads and the only ones the
 * reference requests through Tensor::get); the other blocks of the device copy are zero. */
CCSDT_API int ccsdt_put_dense(ccsdt_ctx* ctx, int tensor, const double* host_dense);
/* Same, without waiting for the copy: host_dense (pinned memory, or the copy is staged) must stay valid and unchanged
 * until the next ccsdt_run / ccsdt_run_tasks returns.  The all-alpha blocks of every tensor travel first, on their own
 * stream, and that run starts with the tasks whose six tiles are all alpha while the other spin patterns still arrive. */
CCSDT_API int ccsdt_put_dense_async(ccsdt_ctx* ctx, int tensor, const double* host_dense);
CCSDT_API int ccsdt_put_block(ccsdt_ctx* ctx, int tensor, const uint32_t bid[4], const double* host_block);
CCSDT_API int ccsdt_set_fetch(ccsdt_ctx* ctx, ccsdt_fetch_fn fn, void* user);
/* drops every fetched block from the HBM block store (blocks given by ccsdt_put_block stay): the next run starts cold,
 * as a fresh CCSD_T_Fused_Driver::execute does */
CCSDT_API int ccsdt_clear_blocks(ccsdt_ctx* ctx);
CCSDT_API int ccsdt_set_synthetic(ccsdt_ctx* ctx, uint64_t seed); /* procedural tensors generated on the device */
/* The three V2 tensors formed on the device from Cholesky vectors: host_chol[N][N][ncv] over all spin orbitals in tile
 * order (occupied first), what ExaChem holds as cholVpr.  Replaces setupV2Tensors (exachem/cholesky/v2tensors.cpp:52-90,
 * called at exachem/cc/ccsd_t/ccsd_t.cpp:168-193): one FP64 DMMA GEMM over the Cholesky index per tensor (hand-written,
 * csrc/ccsdt_v2.cu; no library call) plus an antisymmetrising gather.  Equivalent to ccsdt_put_dense on v2ijab, v2ijka and v2iabc. */
CCSDT_API int ccsdt_put_cholesky(ccsdt_ctx* ctx, const double* host_chol, int64_t ncv);

/* The one collective of the path, inside the C ABI: ncclAllReduce(sum) of {E[T], E(T)} over all ranks, on the context's
 * GPU (replaces the two ec.pg().reduce calls of exachem/cc/ccsd_t/ccsd_t.cpp:262-263).  libnccl.so.2 is loaded on first
 * use.  ccsdt_comm_unique_id fills 128 bytes on one rank (ncclGetUniqueId); the caller broadcasts them (MPI_Bcast, a file,
 * a torch store) and every rank calls ccsdt_comm_init.  After ccsdt_comm_allreduce every rank holds the totals: a caller
 * that reduces again (ExaChem does) must use the value of rank 0 only -- the C++ adapter returns 0 on the other ranks. */
CCSDT_API int ccsdt_comm_unique_id(void* id128);
CCSDT_API int ccsdt_comm_init(ccsdt_ctx* ctx, const void* id128, int rank, int nranks);
CCSDT_API int ccsdt_comm_allreduce(ccsdt_ctx* ctx, double energies[2]);
CCSDT_API int ccsdt_comm_destroy(ccsdt_ctx* ctx);
/* A process-shared int64 task counter in POSIX shared memory (shm_open + mmap) for ccsdt_set_task_counter: the rank with
 * create = 1 makes and zeroes it, the others attach after a barrier; ccsdt_task_counter_close(ptr, name, unlink). */
CCSDT_API int ccsdt_task_counter_open(const char* name, int create, int64_t** counter);
CCSDT_API int ccsdt_task_counter_close(int64_t* counter, const char* name, int unlink_it);

/* Node-shared block store (placement): the ranks of one node (one process per GPU) attach to a directory in
 * POSIX shared memory `name`; a block is then pulled through the fetch callback by ONE rank of the node and copied from that
 * rank's HBM by the others (CUDA IPC mapping, device-to-device over NVLink / NVSwitch) -- host-to-device traffic of the node
 * is about that of a single rank, whatever the hand-out.  The rank with create = 1 makes the segment, the others attach
 * after a barrier; ccsdt_share_detach (also called by ccsdt_destroy) is collective among the attached ranks.  Applies to
 * blocks that arrive through ccsdt_set_fetch. */
CCSDT_API int ccsdt_share_attach(ccsdt_ctx* ctx, const char* name, int local_rank, int local_ranks, int create);
CCSDT_API int ccsdt_share_detach(ccsdt_ctx* ctx);

/* Dynamic task hand-out across ranks: `counter` points to an int64 in memory shared by all ranks of the
 * node (POSIX/SysV shared memory, an MPI shared window, ...), zeroed before every ccsdt_run by one rank
 * with a barrier on both sides.  Each rank then claims tasks of the range in descending-cost order with
 * one atomic fetch-add per task (replaces AtomicCounterGA, ccsd_t_fused_driver.hpp:169-172,456).
 * NULL (default) = the static cost-balanced split selected by options.rank / options.nranks. */
CCSDT_API int ccsdt_set_task_counter(ccsdt_ctx* ctx, int64_t* counter);

/* runs kernel tasks [task_begin, task_end) of the canonical list that belong to this rank
 * (task_end < 0 = to the end).  energies[0] = E[T] partial, energies[1] = E(T) partial, both already
 * weighted by the task factors and summed in task order.  per_task (optional) receives 2 doubles
 * per task of the range (zeros for tasks owned by other ranks). */
CCSDT_API int ccsdt_run(ccsdt_ctx* ctx, int64_t task_begin, int64_t task_end, double energies[2],
              double* per_task, ccsdt_stats* stats);

/* same for an explicit list of task ids (indices into the canonical list); per_task is indexed by list
 * position.  With nranks > 1 and no task counter the list is split among the ranks on its own
 * (longest-processing-time greedy, identical on every rank). */
CCSDT_API int ccsdt_run_tasks(ccsdt_ctx* ctx, const int64_t* task_ids, int64_t n, double energies[2],
                              double* per_task, ccsdt_stats* stats);

/* diagnostics used by tests and bench (device microbenchmarks and unit probes) */
CCSDT_API int ccsdt_probe_fp64_peak(int device, int use_dmma, int iters, double* tflops, double* ms);
CCSDT_API int ccsdt_probe_mainloop(int device, int ta, int tb, int warps_per_cta, int ctas_per_sm, int iters,
                                  double* tflops);
CCSDT_API int ccsdt_probe_dmma_layout(int device, double* c_out /*64*/, const double* a /*8x4*/, const double* b /*4x8*/);
CCSDT_API int ccsdt_probe_tma_swizzle(int device, double* smem_dump /*rows*16*/, int rows);
CCSDT_API int ccsdt_synth_block(int device, uint64_t seed, int tensor, int noa, int nob, int nva, int nvb,
                      const int64_t lo[4], const int64_t n[4], double* host_out);

#ifdef __cplusplus
}
#endif
#endif /* CCSDT_B200_H */
