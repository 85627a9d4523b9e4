// pdl_chain.h -- host-side bookkeeping for the overlapped (programmatic dependent) launches of the conv-stack kernel.
//
// A launch may skip `griddepcontrol.wait` (conv_fused.cuh) only when it is provably independent of every launch it could
// overlap with.  The chain below holds the address ranges of the launches since the last one that DID wait; a waiting launch is
// a full fence for the chain because every member fills all SMs with one CTA each (so a CTA of launch B only starts on an SM
// after the CTAs of A, A-1, ... there have exited: "A complete" implies the whole chain is).  Pure C++, no CUDA: replayed on the
// CPU by tests/test_boundary_cpu.py through cnnacc_pdl_chain_host().
#pragma once
#include <cstdint>

namespace cnnacc {

struct PdlRange { uintptr_t in_lo, in_hi, out_lo, out_hi; };

struct PdlChain {
    static constexpr int kMax = 16;
    PdlRange hist[kMax];
    int n = 0;                      // launches in the chain
    int64_t launch_id = -1;         // the handle's launch counter right after the newest member
    int64_t global_seq = -1;        // the process-wide conv-launch sequence number the newest member got
    uintptr_t stream = 0;

    static bool overlap(uintptr_t a0, uintptr_t a1, uintptr_t b0, uintptr_t b1) { return a0 < b1 && b0 < a1; }

    // Decide for a launch of `n_images` (grid = min(n_images, sm_count)) with ranges `r` on `stream_now`, given the handle's
    // launch counter and the global sequence number as they are NOW (before this launch), then record it.
    // Returns 1 = the kernel must wait for its predecessor, 0 = it may overlap it.
    int decide(const PdlRange& r, int64_t n_images, int sm_count, uintptr_t stream_now, int64_t handle_launches_now,
               int64_t global_seq_now, int64_t global_seq_mine) {
        bool indep = n_images >= sm_count && n > 0 && n < kMax && launch_id == handle_launches_now && stream == stream_now &&
                     global_seq == global_seq_now;
        for (int i = 0; indep && i < n; i++) {
            const PdlRange& p = hist[i];
            indep = !overlap(r.in_lo, r.in_hi, p.out_lo, p.out_hi) &&        // reads what a chain member writes
                    !overlap(r.out_lo, r.out_hi, p.out_lo, p.out_hi) &&      // writes what a chain member writes
                    !overlap(r.out_lo, r.out_hi, p.in_lo, p.in_hi);          // writes what a chain member still reads
        }
        if (!indep) n = 0;                               // this launch waits: everything before it is complete when it runs
        if (n_images >= sm_count) hist[n++] = r;         // a partial grid is no fence for its successors: never a chain member
        stream = stream_now;
        global_seq = global_seq_mine;
        launch_id = handle_launches_now + 1;
        return indep ? 0 : 1;
    }
    void reset() { n = 0; }
};

}  // namespace cnnacc
