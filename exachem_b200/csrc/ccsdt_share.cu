// Node-shared block directory: the ranks of one node (one process per GPU) fetch every tensor block from the host ONCE
// and read each other's copies over NVLink afterwards (SURVEY.md 8e "placement"; the reference does one-sided
// Tensor::get per rank, exachem/cc/ccsd_t/ccsd_t_all_fused_doubles2.hpp:215,230,335, and every rank keeps its own host
// cache, exachem/cc/ccsd_t/ccsd_t.cpp:236-241).
//
//   * A POSIX shared-memory segment holds a hash table  block id -> (owner rank, slab, offset, state, readers)  and, per
//     rank, the CUDA IPC handles of the slabs its block store is carved from (cudaMalloc'ed 256 MiB to 4 GiB at a time:
//     pool allocations of cudaMallocAsync cannot be exported as legacy IPC handles).
//   * A rank that needs a block looks it up under the segment's process-shared mutex.  READY at another rank: it bumps
//     the entry's reader count, maps the owner's slab (cudaIpcOpenMemHandle, once per slab) and copies the block into its
//     own store device-to-device on its fetch stream -- NVLink / NVSwitch, no host involved.  Absent: it inserts the entry
//     as FETCHING, pulls the block through the fetch callback as usual and publishes it (READY) once its host-to-device
//     copy has completed (a host function queued behind the copy on the fetch stream).  FETCHING elsewhere: it waits a
//     few milliseconds for the owner's upload, then copies from it; past 50 ms it fetches a private copy instead.
//   * The owner never evicts a block with readers; a reader drops its count when its copy has completed.
//   * Detach is collective: barrier (nobody is still reading), every rank closes its mappings, barrier, only then the
//     owners free their slabs (freeing exported memory that an importer still has open is undefined), barrier, unmap.
#include "ccsdt_ctx.hpp"

#include <fcntl.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <thread>

using namespace ccsdt;

namespace ccsdt {

constexpr int      kMaxRanks = 16, kMaxSlabs = 256;
constexpr uint32_t kEmpty = 0, kFetching = 1, kReady = 2, kTomb = 3;
constexpr size_t   kSlabBytes = (size_t) 4 << 30;

struct ShareEntry {
  uint64_t key;     // tensor << 60 | bid[0] << 45 | bid[1] << 30 | bid[2] << 15 | bid[3], + 1 (0 = never used)
  uint32_t state;
  int32_t  owner;
  int32_t  slab;
  uint32_t readers;
  uint64_t offset, bytes;
};
struct ShareRank {
  uint32_t           nslabs;
  int32_t            device;
  cudaIpcMemHandle_t handle[kMaxSlabs];
};
struct ShareHeader {
  pthread_mutex_t mutex;
  uint32_t        magic, nranks, capacity, pad_;
  int64_t         barrier_count, barrier_gen;
  ShareRank       ranks[kMaxRanks];
  // ShareEntry table[capacity] follows
};

struct Slab {
  double* base  = nullptr;
  size_t  bytes = 0, used = 0;
};
struct ShareState {
  ShareHeader* hdr = nullptr;
  ShareEntry*  table = nullptr;
  size_t       map_bytes = 0;
  int          rank = 0, nranks = 1;
  std::string  name;
  bool         creator = false;
  std::vector<Slab> slabs;
  std::map<size_t, std::vector<std::pair<int, size_t>>> free_list; // exact-size reuse: block sizes repeat heavily
  std::map<std::pair<int, int>, void*>                  peer_base; // (rank, slab) -> mapped base
  std::vector<ShareEntry*>                              mine;      // directory entries this rank has claimed
};

// runs on a CUDA callback thread when the copy queued before it on the fetch stream has completed: an upload publishes
// its block (FETCHING -> READY, unless the owner dropped it meanwhile), a peer copy gives the owner's block back
struct CopyDone {
  ShareEntry* e;
  bool        publish;
};
static void CUDART_CB copy_done(void* p) {
  auto* d = static_cast<CopyDone*>(p);
  if(d->publish) {
    uint32_t expect = kFetching;
    __atomic_compare_exchange_n(&d->e->state, &expect, kReady, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
  }
  else __atomic_fetch_sub(&d->e->readers, 1u, __ATOMIC_ACQ_REL);
  delete d;
}

static uint64_t pack_key(const BlockKey& k) {
  return (((uint64_t) k.tensor << 60) | ((uint64_t) k.b[0] << 45) | ((uint64_t) k.b[1] << 30) | ((uint64_t) k.b[2] << 15) |
          (uint64_t) k.b[3]) + 1;
}
static uint64_t mix(uint64_t x) {
  x ^= x >> 33, x *= 0xff51afd7ed558ccdull, x ^= x >> 33, x *= 0xc4ceb9fe1a85ec53ull, x ^= x >> 33;
  return x;
}

static void share_lock(ShareState* s) {
  const int rc = pthread_mutex_lock(&s->hdr->mutex);
  if(rc == EOWNERDEAD) pthread_mutex_consistent(&s->hdr->mutex); // a rank died holding it: the table stays usable
}
static void share_unlock(ShareState* s) { pthread_mutex_unlock(&s->hdr->mutex); }

static void share_barrier(ShareState* s) {
  int64_t *     count = &s->hdr->barrier_count, *gen = &s->hdr->barrier_gen;
  const int64_t g = __atomic_load_n(gen, __ATOMIC_ACQUIRE);
  if(__atomic_add_fetch(count, 1, __ATOMIC_ACQ_REL) == s->nranks) {
    __atomic_store_n(count, 0, __ATOMIC_RELAXED);
    __atomic_add_fetch(gen, 1, __ATOMIC_ACQ_REL);
    return;
  }
  const auto t0 = std::chrono::steady_clock::now();
  while(__atomic_load_n(gen, __ATOMIC_ACQUIRE) == g) {
    std::this_thread::sleep_for(std::chrono::microseconds(50));
    if(std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return; // a rank is gone: do not hang the others
  }
}

// looks the key up; with `insert` claims a free / tombstoned slot for it when absent (caller holds the lock)
static ShareEntry* share_find(ShareState* s, uint64_t key, bool insert) {
  const uint32_t cap = s->hdr->capacity;
  ShareEntry*    tomb = nullptr;
  for(uint32_t i = (uint32_t) (mix(key) % cap), n = 0; n < cap; n++, i = (i + 1) % cap) {
    ShareEntry& e = s->table[i];
    if(e.key == key && e.state != kTomb) return &e;
    if(e.key == key && e.state == kTomb && !tomb && e.readers == 0) tomb = &e;
    if(e.key == 0) {
      if(!insert) return nullptr;
      ShareEntry* slot = tomb ? tomb : &e;
      slot->key        = key;
      slot->state      = kEmpty;
      slot->readers    = 0;
      return slot;
    }
    if(e.state == kTomb && !tomb && e.readers == 0) tomb = &e;
  }
  if(insert && tomb) {
    tomb->key = key, tomb->state = kEmpty, tomb->readers = 0;
    return tomb;
  }
  return nullptr;
}

int share_alloc(ccsdt_ctx* ctx, size_t bytes, double** dev, int* slab, size_t* offset) {
  ShareState* s = ctx->share;
  bytes         = (bytes + 255) & ~(size_t) 255;
  auto fl       = s->free_list.find(bytes);
  if(fl != s->free_list.end() && !fl->second.empty()) {
    *slab   = fl->second.back().first;
    *offset = fl->second.back().second;
    fl->second.pop_back();
    *dev = (double*) ((char*) s->slabs[*slab].base + *offset);
    return 0;
  }
  for(int pass = 0; pass < 2; pass++) {
    if(!s->slabs.empty() && s->slabs.back().used + bytes <= s->slabs.back().bytes) {
      Slab& sl = s->slabs.back();
      *slab    = (int) s->slabs.size() - 1;
      *offset  = sl.used;
      *dev     = (double*) ((char*) sl.base + sl.used);
      sl.used += bytes;
      return 0;
    }
    if((int) s->slabs.size() >= kMaxSlabs) return ctx->fail("shared block store: slab table full", 2);
    Slab sl;
    // 256 MiB to start with, doubling up to 4 GiB: few slabs (each costs its peers one cudaIpcOpenMemHandle) without
    // reserving gigabytes for a job of megabytes
    sl.bytes = std::max(std::min<size_t>((size_t) 256 << 20 << std::min<size_t>(s->slabs.size(), 4), kSlabBytes), bytes);
    cudaError_t e = cudaMalloc((void**) &sl.base, sl.bytes);
    if(e != cudaSuccess) {
      cudaGetLastError();
      return -1; // out of memory: the caller evicts and retries
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, sl.base));
    ShareRank& me                 = s->hdr->ranks[s->rank];
    me.handle[s->slabs.size()]    = h;
    __atomic_store_n(&me.nslabs, (uint32_t) s->slabs.size() + 1, __ATOMIC_RELEASE);
    s->slabs.push_back(sl);
  }
  return ctx->fail("shared block store: allocation failed", 2);
}

size_t share_unused_bytes(const ccsdt_ctx* ctx) {
  size_t n = 0;
  if(!ctx->share) return 0;
  for(const Slab& sl: ctx->share->slabs) n += sl.bytes - sl.used;
  for(const auto& kv: ctx->share->free_list) n += kv.first * kv.second.size();
  return n;
}

void share_free(ccsdt_ctx* ctx, int slab, size_t offset, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t) 255;
  ctx->share->free_list[bytes].push_back({slab, offset});
}

// every copy this rank has queued on its fetch stream has completed: its uploads are published, the blocks it copied from
// peers are given back
void share_poll(ccsdt_ctx* ctx, bool wait) {
  if(ctx->share && wait && ctx->s_fetch) cudaStreamSynchronize(ctx->s_fetch);
}

// A block that is not in this rank's store.  Returns 1 when a peer copy was issued into *dev (the block is on its way),
// 0 when the caller must fetch it from the host into *dev -- *entry then is the directory entry to publish after the
// upload (NULL: private copy) -- and < 0 on out-of-memory (evict and retry) or > 1 on error.
int share_acquire(ccsdt_ctx* ctx, const BlockKey& key, size_t bytes, double** dev, int* slab, size_t* offset, void** entry) {
  ShareState* s = ctx->share;
  *entry = nullptr;
  if(int rc = share_alloc(ctx, bytes, dev, slab, offset)) return rc < 0 ? -1 : 2;
  const uint64_t k = pack_key(key);
  const auto     t0 = std::chrono::steady_clock::now();
  for(;;) {
    share_lock(s);
    ShareEntry* e = share_find(s, k, true);
    if(e && e->state == kFetching && e->owner != s->rank &&
       std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(50)) {
      share_unlock(s); // another rank is uploading it right now: cheaper to wait for that than to fetch it again
      std::this_thread::sleep_for(std::chrono::microseconds(20));
      continue;
    }
    if(e && e->state == kReady && e->owner != s->rank && e->bytes == bytes) {
      e->readers++;
      const int    owner = e->owner, oslab = e->slab;
      const size_t ooff = e->offset;
      share_unlock(s);
      auto pb = s->peer_base.find({owner, oslab});
      if(pb == s->peer_base.end()) {
        void*       base = nullptr;
        cudaError_t ce   = cudaIpcOpenMemHandle(&base, s->hdr->ranks[owner].handle[oslab], cudaIpcMemLazyEnablePeerAccess);
        if(ce != cudaSuccess) {
          cudaGetLastError();
          __atomic_fetch_sub(&e->readers, 1u, __ATOMIC_ACQ_REL);
          return 0; // no peer path between the two devices: fetch from the host, privately
        }
        pb = s->peer_base.emplace(std::make_pair(owner, oslab), base).first;
      }
      cudaError_t ce = cudaMemcpyAsync(*dev, (const char*) pb->second + ooff, bytes, cudaMemcpyDefault, ctx->s_fetch);
      if(ce == cudaSuccess) ce = cudaLaunchHostFunc(ctx->s_fetch, copy_done, new CopyDone{e, false});
      if(ce != cudaSuccess) {
        __atomic_fetch_sub(&e->readers, 1u, __ATOMIC_ACQ_REL);
        return ctx->fail(std::string("peer copy failed: ") + cudaGetErrorString(ce), 2);
      }
      ctx->fetch_dirty = true;
      ctx->stats.blocks_from_peers++;
      ctx->stats.peer_bytes += (int64_t) bytes;
      return 1;
    }
    if(e && e->state == kEmpty) { // absent: this rank fetches it for the node
      e->state  = kFetching;
      e->owner  = s->rank;
      e->slab   = *slab;
      e->offset = *offset;
      e->bytes  = bytes;
      *entry    = e;
      s->mine.push_back(e);
    }
    share_unlock(s);
    return 0;
  }
}

// the upload of a block this rank fetches for the node has been issued: publish it when `ev`'s copy has landed
int share_publish_after(ccsdt_ctx* ctx, void* entry, cudaStream_t st) {
  if(!entry) return 0;
  CK(cudaLaunchHostFunc(st, copy_done, new CopyDone{(ShareEntry*) entry, true}));
  return 0;
}

// this rank drops a block of its shared store: false = a peer is still copying it (try another victim)
bool share_release(ccsdt_ctx* ctx, const BlockKey& key, BlockEntry& be, bool wait) {
  ShareState* s = ctx->share;
  share_lock(s);
  ShareEntry* e = share_find(s, pack_key(key), false);
  if(e && e->owner == s->rank && e->slab == be.slab && e->offset == be.offset) {
    while(e->readers > 0) {
      share_unlock(s);
      if(!wait) return false;
      std::this_thread::sleep_for(std::chrono::microseconds(50));
      share_lock(s);
    }
    e->state = kTomb; // (an upload still in flight finds the entry no longer FETCHING and does not publish it)
  }
  share_unlock(s);
  share_free(ctx, be.slab, be.offset, be.bytes);
  return true;
}

void share_detach(ccsdt_ctx* ctx) {
  ShareState* s = ctx->share;
  if(!s) return;
  cudaSetDevice(ctx->device);
  if(ctx->s_fetch) cudaStreamSynchronize(ctx->s_fetch);
  share_poll(ctx, true);
  share_barrier(s); // nobody reads anybody's slabs any more
  for(auto& kv: s->peer_base) cudaIpcCloseMemHandle(kv.second);
  share_barrier(s); // every importer has closed its mappings: only now may the exporters free (CUDA IPC rule)
  // every block of the shared store goes with its slab
  for(auto it = ctx->blocks.begin(); it != ctx->blocks.end();) {
    if(it->second.slab >= 0) {
      ctx->block_bytes -= it->second.bytes;
      it = ctx->blocks.erase(it);
    }
    else ++it;
  }
  share_lock(s);
  for(ShareEntry* e: s->mine)
    if(e->owner == s->rank && e->state != kTomb) e->state = kTomb; // (an entry another rank has taken over since is not ours)
  s->hdr->ranks[s->rank].nslabs = 0;
  share_unlock(s);
  if(ctx->s_stage) cudaStreamSynchronize(ctx->s_stage);
  for(Slab& sl: s->slabs) cudaFree(sl.base);
  share_barrier(s); // every rank has let go of the segment's contents
  munmap(s->hdr, s->map_bytes);
  if(s->creator) shm_unlink(s->name.c_str());
  delete s;
  ctx->share = nullptr;
}

} // namespace ccsdt

extern "C" {

int ccsdt_share_attach(ccsdt_ctx* ctx, const char* name, int local_rank, int local_ranks, int create) {
  if(!ctx || !name || local_ranks < 1 || local_ranks > kMaxRanks || local_rank < 0 || local_rank >= local_ranks) return 1;
  if(ctx->share) return ctx->fail("a shared block directory is already attached");
  cudaSetDevice(ctx->device);
  clear_blocks(ctx, true); // blocks of the private pool are not visible to peers: start clean
  const uint32_t capacity = 1u << 20;
  const size_t   bytes    = sizeof(ShareHeader) + sizeof(ShareEntry) * (size_t) capacity;
  if(create) shm_unlink(name);
  const int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
  if(fd < 0) return ctx->fail(std::string("shm_open(") + name + ") failed", 11);
  if(create && ftruncate(fd, (off_t) bytes) != 0) {
    close(fd);
    shm_unlink(name);
    return ctx->fail("ftruncate of the shared block directory failed", 11);
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if(p == MAP_FAILED) return ctx->fail("mmap of the shared block directory failed", 11);
  auto* s      = new ShareState;
  s->hdr       = (ShareHeader*) p;
  s->table     = (ShareEntry*) ((char*) p + sizeof(ShareHeader));
  s->map_bytes = bytes;
  s->rank      = local_rank;
  s->nranks    = local_ranks;
  s->name      = name;
  s->creator   = create != 0;
  if(create) {
    // the segment is zero-filled by ftruncate: every table entry is empty
    pthread_mutexattr_t at;
    pthread_mutexattr_init(&at);
    pthread_mutexattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_mutexattr_setrobust(&at, PTHREAD_MUTEX_ROBUST);
    pthread_mutex_init(&s->hdr->mutex, &at);
    pthread_mutexattr_destroy(&at);
    s->hdr->nranks   = (uint32_t) local_ranks;
    s->hdr->capacity = capacity;
    __atomic_store_n(&s->hdr->magic, 0xCC5D7B20u, __ATOMIC_RELEASE);
  }
  else if(__atomic_load_n(&s->hdr->magic, __ATOMIC_ACQUIRE) != 0xCC5D7B20u || s->hdr->nranks != (uint32_t) local_ranks) {
    munmap(p, bytes);
    delete s;
    return ctx->fail("shared block directory is not initialised (attach after the creating rank, behind a barrier)", 11);
  }
  s->hdr->ranks[local_rank].device = ctx->device;
  s->hdr->ranks[local_rank].nslabs = 0;
  ctx->share                       = s;
  return 0;
}

int ccsdt_share_detach(ccsdt_ctx* ctx) {
  if(!ctx) return 1;
  share_detach(ctx);
  return 0;
}

} // extern "C"
