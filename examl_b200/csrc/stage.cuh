// stage.cuh -- entry windows: how the traverse kernels bring the descriptor and the tip codes on chip.
//
// A traversal descriptor can be long (one entry per inner node: tens of thousands on the trees ExaML is used for)
// and every entry may touch one or two tip rows, so neither fits in shared memory as a whole.  The kernels therefore
// walk the descriptor in WINDOWS of kWindow entries.  Per window and per site chunk
//   1. all threads expand the window's entries (node numbers -> CLV element offsets / tip-row slots) into smem,
//   2. warp 0 issues one cp.async.bulk (TMA, 1-D; SASS UBLKCP) per tip row the window touches -- the chunk's
//      slice of that taxon's tip codes -- all completing on one mbarrier,
//   3. everybody waits on the mbarrier, then the window's entries are processed without any further barrier.
// The cold HBM reads of tip codes are thus issued up front and in bulk; the carry registers live across windows.
// Shared memory is bounded by the window, not by the tree: (kWindow + 1) entries + (2 kWindow + 2) tip rows.
#pragma once
#include "engine.cuh"

namespace exb {

constexpr int kWindow = 96;                  // entries per window

struct SevView {           // what stage_window needs to resolve (node, tile) pairs under -S; table == null: dense arena
  const double *const *table; const double *gapcol; const double *clv;
  int ntiles, tile, part_cls, nparts_cls, span; unsigned tile_site0;
};

struct SEntry {            // a traversal entry as the threads consume it (32 B)
  int tipCase;             // -1 for the branch entry
  int flags;               // kFlag*: which operand is a tip (branch entry) / sits in the carry registers
  long long poff, qoff, roff;   // inner node: element offset of its CLV (node_index * N * span, unique per node);
                                // tip: byte offset of its staged row inside the window's tip area
};

// branch entry: p / q is a tip.  Any entry: the operand is the CLV the previous entry of this call produced, i.e. it is
// still in the thread's carry registers (newview entries: q / r; branch entry: p / q).
constexpr int kFlagPTip = 1, kFlagQTip = 2, kFlagCarryA = 4, kFlagCarryB = 8;
// -S: the operand (newview: q / r, branch: p / q) resp. the newview entry's parent is an all-gap (node, tile): the operand
// is read from the node's gap column (its offset then lacks the per-site term), the parent is neither computed nor stored
constexpr int kFlagGapA = 16, kFlagGapB = 32, kFlagGapP = 64;
// exb_score_insertions: the entry forms the CLV of a candidate's new node in registers and scores it against the pruned
// subtree (the branch entry's p) instead of storing it; poff then holds the candidate index
constexpr int kFlagScore = 128;

__device__ __forceinline__ uint32_t stage_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(stage_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(stage_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(stage_smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(stage_smem_u32(dst)), "l"(src), "r"(bytes), "r"(stage_smem_u32(bar)) : "memory");
}

// Shared-memory layout of a window of KW entries with TSB-byte tip rows:
//   [KW + 1] SEntry | [2 KW + 2][TSB] tip rows | (optional) [KW + 1][DROW] doubles: the entries' d-table rows
template <int TSB, int KW = kWindow>          // TSB: bytes per staged tip row = sites per chunk
__host__ __device__ constexpr size_t window_smem_bytes(int drow_doubles = 0)
{
  return (size_t)(KW + 1) * sizeof(SEntry) + (size_t)(2 * KW + 2) * TSB + (size_t)(KW + 1) * drow_doubles * sizeof(double);
}
template <int TSB, int KW = kWindow>
__host__ __device__ constexpr size_t window_dtab_offset()
{
  return (size_t)(KW + 1) * sizeof(SEntry) + (size_t)(2 * KW + 2) * TSB;
}

// Stage entries [w0, w1) (slots 0 .. w1-w0-1) and, when `with_branch`, the branch entry 0 (slot KW).
// Must be called by every thread of the CTA; ends with all staged data visible to all threads.  Returns the mbarrier
// phase for the next call.  Takes plain values (not the kernel's argument struct) so that nothing of the caller has to
// live in local memory for this non-inlined call.
//   dyn      : dynamic shared memory base (16-byte aligned)
//   span_el  : doubles per site in a CLV (node stride = N * span_el)
//   first, count : the call's newview entries are [first, count) -- carry flags look at the predecessor entry
//   drow     : null, or this partition's row of entry 0 in the d-table; rows of `drow_doubles` doubles, `dstride` doubles
//              apart, are staged behind the tips (row of entry e in slot e - w0, the branch entry's row 0 in slot KW): the
//              exponentials of an entry then cost the threads a shared-memory read instead of an L2 round trip
//   scratch_off : element offset (relative to the CLV arena) of scratch node 2*mxtips-1, the first of the nodes that only
//              exb_score_insertions uses
//   sev      : -S view of this tile (table == null: dense arena).  Inner-node offsets then come from the block table
//              (relative to `clv`, minus the tile's first site so that the kernels' site term still applies), all-gap
//              operands point at the node's gap column and are flagged
template <int TSB, int KW = kWindow>
__device__ __noinline__ uint32_t stage_window(unsigned char *dyn, uint64_t *mbar, uint32_t phase, const DevEntry *entries,
                                              const uint8_t *tips, int mxtips, size_t N, int span_el, size_t chunk_site0,
                                              unsigned nchunk, int w0, int w1, int first, int count, bool with_branch,
                                              const double *drow = nullptr, size_t dstride = 0, int drow_doubles = 0,
                                              SevView sev = SevView{nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0u},
                                              long long scratch_off = 0)
{
  SEntry *ents = reinterpret_cast<SEntry *>(dyn);
  uint8_t *stips = dyn + (size_t)(KW + 1) * sizeof(SEntry);
  double *sdtab = reinterpret_cast<double *>(dyn + window_dtab_offset<TSB, KW>());
  const int tid = threadIdx.x, lane = tid & 31;
  __syncthreads();                             // the previous window's readers are done with ents / stips
  for (int j = tid; j < (w1 - w0) + (with_branch ? 1 : 0); j += blockDim.x) {
    const bool is_branch = j == (w1 - w0);
    const DevEntry en = entries[is_branch ? 0 : w0 + j];
    const int slot = is_branch ? KW : j;
    // the node whose CLV the threads still hold when this entry is processed: the previous newview entry's parent
    // (a scoring entry leaves a CLV that belongs to no node: its p is negative and matches nothing)
    const int prev = is_branch ? (count > first ? entries[count - 1].p : -1) : (w0 + j > first ? entries[w0 + j - 1].p : -1);
    const bool score = !is_branch && (en.tipCase & kScoreBit);
    SEntry se;
    se.tipCase = is_branch ? -1 : (en.tipCase & 3);
    if (is_branch) se.flags = (en.p <= mxtips ? kFlagPTip : 0) | (en.q <= mxtips ? kFlagQTip : 0) |
                              (en.p == prev ? kFlagCarryA : 0) | (en.q == prev ? kFlagCarryB : 0);
    else           se.flags = (en.q == prev ? kFlagCarryA : 0) | (en.r == prev ? kFlagCarryB : 0) | (score ? kFlagScore : 0);
    // inner nodes: CLV offset in the arena (scratch nodes: in their own allocation); tips: q uses row 2*slot, r (or a
    // tip p of the branch entry) row 2*slot+1
    auto node_off = [&](int node, int tiprow) -> long long {
      if (node <= mxtips) return (long long)tiprow * TSB;
      if (node > 2 * mxtips - 2) return scratch_off + (long long)(node - (2 * mxtips - 1)) * (long long)N * span_el;
      return (long long)(node - mxtips - 1) * (long long)N * span_el;
    };
    se.poff = score ? (long long)(-en.p - 1) : node_off(en.p, 2 * slot + 1);
    se.qoff = node_off(en.q, 2 * slot);
    se.roff = node_off(en.r, 2 * slot + 1);
    if (sev.table) {
      // inner nodes: block of (node, tile), or the node's gap column when the pair is all-gap
      auto resolve = [&](int node, long long &off, int gapflag, int carryflag) {
        if (node <= mxtips || node == 0) return;
        const double *blk = sev.table[(size_t)(node - mxtips - 1) * sev.ntiles + sev.tile];
        if (blk) off = (long long)(blk - sev.clv) - (long long)sev.tile_site0 * sev.span;
        else {
          off = (long long)(sev.gapcol + ((size_t)(node - mxtips - 1) * sev.nparts_cls + sev.part_cls) * sev.span - sev.clv);
          se.flags = (se.flags | gapflag) & ~carryflag;
        }
      };
      if (is_branch) { resolve(en.p, se.poff, kFlagGapA, kFlagCarryA); resolve(en.q, se.qoff, kFlagGapB, kFlagCarryB); }
      else { resolve(en.p, se.poff, kFlagGapP, 0); resolve(en.q, se.qoff, kFlagGapA, kFlagCarryA); resolve(en.r, se.roff, kFlagGapB, kFlagCarryB); }
    }
    ents[slot] = se;
  }
  __syncthreads();
  if (tid < 32) {
    const uint32_t bytes = (nchunk + 15u) & ~15u;
    const uint32_t dbytes = (uint32_t)drow_doubles * (uint32_t)sizeof(double);
    const uint8_t *gsrc = tips + chunk_site0;
    const int n = (w1 - w0) + (with_branch ? 1 : 0);
    int mine = 0;
    for (int j = lane; j < n; j += 32) {
      const bool is_branch = j == (w1 - w0);
      const DevEntry en = entries[is_branch ? 0 : w0 + j];
      if (is_branch) mine += (en.p <= mxtips) + (en.q <= mxtips);
      else           mine += ((en.tipCase & 3) <= 1) + ((en.tipCase & 3) == 0);
    }
    const int total = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) mbar_expect_tx(mbar, (uint32_t)total * bytes + (drow ? (uint32_t)n * dbytes : 0u));
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
      const bool is_branch = j == (w1 - w0);
      const DevEntry en = entries[is_branch ? 0 : w0 + j];
      const int slot = is_branch ? KW : j;
      if (is_branch) {
        if (en.p <= mxtips) bulk_g2s(stips + (size_t)(2 * slot + 1) * TSB, gsrc + (size_t)(en.p - 1) * N, bytes, mbar);
        if (en.q <= mxtips) bulk_g2s(stips + (size_t)(2 * slot) * TSB, gsrc + (size_t)(en.q - 1) * N, bytes, mbar);
      } else {
        if ((en.tipCase & 3) <= 1) bulk_g2s(stips + (size_t)(2 * slot) * TSB, gsrc + (size_t)(en.q - 1) * N, bytes, mbar);
        if ((en.tipCase & 3) == 0) bulk_g2s(stips + (size_t)(2 * slot + 1) * TSB, gsrc + (size_t)(en.r - 1) * N, bytes, mbar);
      }
      if (drow) bulk_g2s(sdtab + (size_t)slot * drow_doubles, drow + (size_t)(is_branch ? 0 : w0 + j) * dstride, dbytes, mbar);
    }
  }
  mbar_wait(mbar, phase);
  return phase ^ 1u;
}

}  // namespace exb
