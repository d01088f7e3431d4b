// Weight packing as a handful of launches.  selfc_ctx_load_weights packs ~300 weight tensors into the images the kernels read; a
// training step does that after every optimiser step, and one tiny launch (+ one 128-byte copy) per tensor was 7 % of the step.  While a
// PackBatch is current (pack_batch_current()), launch_pack_conv_simt / pack_tc_weights / pack_temporal_weights only RECORD a job; the
// flush runs every job of a kind in ONE launch (a block looks its job up in a table of first-block offsets).  The job tables live in
// device memory owned by the context and are uploaded only when they differ from what is there (step after step they do not).
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include <vector>

namespace selfc {

struct JobTable {
  std::vector<char> host;        // the job structs recorded since the last flush
  std::vector<int> first;        // first block of each job; first[njobs] = number of blocks
  std::vector<char> uploaded;    // what the device copy holds: [jobs][first]
  char* dev = nullptr;
  size_t dev_bytes = 0;
  size_t job_size = 0;

  int njobs() const { return (int)first.size() - 1; }
  void clear() {
    host.clear();
    first.assign(1, 0);
  }
  template <typename J>
  void add(const J& j, int nblocks) {
    if (first.empty()) first.assign(1, 0);
    job_size = sizeof(J);
    const size_t o = host.size();
    host.resize(o + sizeof(J));
    memcpy(host.data() + o, &j, sizeof(J));
    first.push_back(first.back() + nblocks);
  }
  // device copy of the recorded jobs -> *jobs, *firsts (valid for launches on `st` after this call)
  cudaError_t sync(cudaStream_t st, const void** jobs, const int** firsts) {
    const size_t jb = (host.size() + 15) & ~(size_t)15, fb = first.size() * sizeof(int);
    std::vector<char> img(jb + fb, 0);
    memcpy(img.data(), host.data(), host.size());
    memcpy(img.data() + jb, first.data(), fb);
    if (dev_bytes < img.size()) {
      if (dev) cudaFree(dev);
      dev = nullptr;
      dev_bytes = 0;
      uploaded.clear();
      const size_t cap = img.size() * 2 + 4096;
      cudaError_t e = cudaMalloc(&dev, cap);
      if (e != cudaSuccess) return e;
      dev_bytes = cap;
    }
    if (uploaded != img) {
      // (pageable source: the call returns once the bytes are staged; stream order protects launches still reading the old table)
      cudaError_t e = cudaMemcpyAsync(dev, img.data(), img.size(), cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
      uploaded.swap(img);
    }
    *jobs = dev;
    *firsts = reinterpret_cast<const int*>(dev + jb);
    return cudaSuccess;
  }
  void release() {
    if (dev) cudaFree(dev);
    dev = nullptr;
    dev_bytes = 0;
    uploaded.clear();
  }
};

// One table per (kind, flush point of selfc_ctx_load_weights): a table that holds the same jobs call after call is never re-uploaded.
constexpr int kPackFlushPoints = 4;
struct PackBatch {
  // fp32 packs; conv3x3 images; the training step's input-gradient images (slot images of conv1..4, conv5's flipped weights); temporal
  // / pointwise images
  JobTable simt[kPackFlushPoints], tc3[kPackFlushPoints], slot[kPackFlushPoints], ref5[kPackFlushPoints], temporal[kPackFlushPoints];
  int point = 0;
  void release() {
    for (int i = 0; i < kPackFlushPoints; ++i) {
      simt[i].release();
      tc3[i].release();
      slot[i].release();
      ref5[i].release();
      temporal[i].release();
    }
  }
};

// the batch the calling thread records into (null: every pack launches at once)
PackBatch*& pack_batch_current();
// run what was recorded since the last flush -- fp32 packs first, then the images that read them or the parameters, the temporal /
// pointwise images last (they may read an fp32 pack or conv5's flipped weights) -- and move on to the next set of tables
int pack_batch_flush(PackBatch& b, cudaStream_t st);

// records the packs issued while it lives; whatever path the caller leaves by, later packs launch directly again
struct PackBatchScope {
  explicit PackBatchScope(PackBatch& b) {
    b.point = 0;
    for (int i = 0; i < kPackFlushPoints; ++i) {
      b.simt[i].clear();
      b.tc3[i].clear();
      b.slot[i].clear();
      b.ref5[i].clear();
      b.temporal[i].clear();
    }
    pack_batch_current() = &b;
  }
  ~PackBatchScope() { pack_batch_current() = nullptr; }
  PackBatchScope(const PackBatchScope&) = delete;
  PackBatchScope& operator=(const PackBatchScope&) = delete;
};
bool pack_batch_enabled();      // SELFC_PACK_BATCH=0: one launch per weight tensor (A/B)

// implemented next to the kernels
int flush_pack_conv_simt(JobTable& t, cudaStream_t st);
int flush_pack_tc3(JobTable& t, cudaStream_t st);
int flush_pack_dgrad_slot(JobTable& t, cudaStream_t st);
int flush_pack_dgrad5_ref(JobTable& t, cudaStream_t st);
int flush_pack_temporal(JobTable& t, cudaStream_t st);

// device side: the job of this block.  first[] is ascending with first[0] = 0; returns j with first[j] <= block < first[j + 1]
__device__ __forceinline__ int pack_find_job(const int* __restrict__ first, int njobs, int block) {
  int lo = 0, hi = njobs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(first + mid) <= block) lo = mid; else hi = mid;
  }
  return lo;
}

}  // namespace selfc
