// Host side of hint generation: job validation, launch shaping (CTA width, rounds, shared last round, round barrier,
// PRF variant) and the pm_hintgen / pm_hintgen_dev entry points.  The kernel lives in pm_hintgen.cuh / pm_hg_*.cu.
#include <algorithm>
#include <cstring>
#include <vector>

#include "pm_hg_params.cuh"

namespace pm {

// Enqueue hint generation for `jobs` (all pointers inside are device pointers) on `st`.
int hintgen_enqueue(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs, cudaStream_t st) {
    const uint64_t E = db->entry_u64;
    const bool wide = (E % 2 == 0);  // 16-byte vectors need 16-byte aligned rows
    const uint32_t ev = (uint32_t)(wide ? E / 2 : E), evx = (uint32_t)(wide ? (E & ~3ull) / 2 : (E & ~3ull));
    bool need4 = false;
    for (uint64_t a = 0; a < n_jobs; a++) {
        const pm_hint_job &J = jobs[a];
        if (J.chunk_size == 0 || (J.chunk_size & (J.chunk_size - 1)))
            return set_error(PM_ERR_ARG, "hintgen: chunk_size %llu is not a power of two", (unsigned long long)J.chunk_size);
        if (J.row0 + J.n_rows > db->n_rows) return set_error(PM_ERR_ARG, "hintgen: job %llu exceeds the table", (unsigned long long)a);
        if (J.n_rows >= 0x7fffffffull || J.set_size >= 0x7fffffffull || J.chunk_size > 0x80000000ull ||
            J.chunk_size * J.set_size > 0xffffffffull)
            return set_error(PM_ERR_UNSUPPORTED, "hintgen: instance too large for 32-bit row offsets");
        if (J.n_hints && !J.parity_out) return set_error(PM_ERR_ARG, "hintgen: parity_out is null");
        if (J.chunk_size > 65536) need4 = true;
    }
    int G, nv;
    hg_shape(evx, &G, &nv);
    if (nv > 8) return set_error(PM_ERR_UNSUPPORTED, "hintgen: entry_u64 too large");
    const uint32_t hpw = 32 / G;   // hints per warp
    const uint32_t sm = (uint32_t)db->sm_count;
    uint64_t a = 0;
    while (a < n_jobs) {
        HintParams P;
        memset(&P, 0, sizeof(P));
        P.db = db->d_rows;
        P.ev = ev;
        P.evx = evx;
        // the jobs of this launch
        uint64_t b = a, nj2 = 0, max_slice = 0, max_set = 0;
        for (; b < n_jobs && nj2 < HG_MAX_JOBS; b++) {
            if (jobs[b].n_hints == 0 || jobs[b].n_rows == 0) continue;
            nj2++;
            max_slice = std::max<uint64_t>(max_slice, jobs[b].n_rows * E * 8);
            max_set = std::max<uint64_t>(max_set, jobs[b].set_size);
        }
        // Shared last round (stream-K): on unless the hg_tail_split knob says 0.
        const bool tail_share = tune(T_HG_TAIL_SPLIT) != 0;
        // CTA width.  Every lane group keeps one hint for a whole sweep, so a tile costs one sweep whatever it holds.  A
        // round gets cheaper with fewer warps, but less than proportionally (measured on B200, MS-MARCO rows: round time
        // ~ warps + 24): with the shared last round the cost is proportional to tiles * (warps + 24) and the widest CTA
        // wins; without it a narrower CTA pays when it saves a whole round (1/8 of the hints: 5.2 -> 6 rounds at 16 warps,
        // 5.95 -> 6 at 14).  The hg_warps knob forces a width.
        uint32_t warps = HG_THREADS / 32;
        if (!tail_share) {
            uint64_t best = ~0ull;
            for (uint32_t w = 12; w <= (uint32_t)HG_MAX_THREADS / 32; w++) {
                uint64_t t = 0;
                for (uint64_t c = a; c < b; c++)
                    if (jobs[c].n_hints && jobs[c].n_rows) t += (jobs[c].n_hints + (uint64_t)w * hpw - 1) / ((uint64_t)w * hpw);
                const uint64_t g = std::max<uint64_t>(1, std::min<uint64_t>(t, sm));
                const uint64_t cost = ((t + g - 1) / g) * (w + 24);
                if (cost <= best) { best = cost; warps = w; }
            }
        }
        const int force_warps = tune(T_HG_WARPS);
        if (force_warps >= 1 && force_warps <= HG_MAX_THREADS / 32) warps = (uint32_t)force_warps;
        const uint32_t hpt = hpw * warps;
        P.threads = warps * 32;
        uint32_t tiles = 0, nj = 0;
        for (; a < b; a++) {
            const pm_hint_job &J = jobs[a];
            if (J.n_hints == 0) continue;
            if (J.n_rows == 0) {  // an empty instance has all-zero parities
                PM_CUDA(cudaMemsetAsync(J.parity_out, 0, J.n_hints * E * 8, st));
                continue;
            }
            HintJobDev &D = P.jobs[nj++];
            memcpy(D.rk, J.rk, sizeof(D.rk));
            D.row0 = J.row0; D.n_rows = J.n_rows;
            D.hint_begin = J.hint_begin; D.n_hints = J.n_hints; D.n_primary = J.n_primary; D.backup_group = J.backup_group;
            D.tags = J.tags; D.skip = J.skip_chunk; D.out = J.parity_out;
            D.off = J.chunk_size <= 65536 ? J.offsets_out : nullptr;
            D.chunk_mask = (uint32_t)(J.chunk_size - 1);
            D.chunk_shift = (uint32_t)__builtin_ctzll(J.chunk_size);
            D.set_size = (uint32_t)J.set_size;
            D.tile_begin = tiles;
            tiles += (uint32_t)((J.n_hints + hpt - 1) / hpt);
        }
        if (nj == 0) continue;
        P.n_jobs = nj;
        P.n_tiles = tiles;
        uint32_t grid;
        if (tail_share) {
            grid = sm;
            P.full_rounds = tiles / grid;
            P.tail_tiles = tiles % grid;
            // the slices of the shared round are XORed into the output: zero it first (stream order)
            const uint32_t t0 = P.full_rounds * grid;
            for (uint32_t j = 0; j < nj && P.tail_tiles; j++) {
                const HintJobDev &D = P.jobs[j];
                const uint32_t t_end = j + 1 < nj ? P.jobs[j + 1].tile_begin : tiles;
                if (t_end <= t0) continue;
                const uint64_t h0 = (uint64_t)(std::max(t0, D.tile_begin) - D.tile_begin) * hpt;
                if (h0 < D.n_hints) PM_CUDA(cudaMemsetAsync(D.out + h0 * E, 0, (D.n_hints - h0) * E * 8, st));
            }
        } else {
            grid = std::min(tiles, sm);
            P.full_rounds = (tiles + grid - 1) / grid;
            P.tail_tiles = 0;
        }
        P.serpentine = tune(T_HG_SERPENTINE) != 0;
        // The round barrier pays when a sub-PIR's slice does not fit in L2 (measured: 896 B x 200 k rows = 179 MB,
        // -12 % time); for slices that stay L2-resident anyway (SIFT-shaped: 40 MB) it only costs.  hg_sync forces.
        const int force_sync = tune(T_HG_SYNC);
        const bool use_sync = (force_sync >= 0 ? force_sync != 0 : max_slice > (64ull << 20)) && P.full_rounds + (P.tail_tiles ? 1 : 0) > 1;
        P.sync = use_sync ? sync_counter(db) : nullptr;
        // PRF variant: chunk-id bytes that can vary (rounds 1-2 are hoisted for the others), PRF output bytes kept
        int xb = max_set <= 256 ? 1 : max_set <= 65536 ? 2 : 4;
        const int force_xb = tune(T_HG_XBYTES);
        if (force_xb == 2 || force_xb == 4) xb = std::max(xb, force_xb);
        int rc;
        if (wide) rc = (need4 || xb == 4) ? hg_launch_wide_xb4(P, grid, st) : xb == 2 ? hg_launch_wide_xb2(P, grid, st) : hg_launch_wide_xb1(P, grid, st);
        else rc = (need4 || xb == 4) ? hg_launch_narrow_xb4(P, grid, st) : hg_launch_narrow_xb2(P, grid, st);
        if (rc != PM_OK) return rc;
    }
    return PM_OK;
}

}  // namespace pm

using namespace pm;

PM_EXPORT int pm_hintgen_dev(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs, void *stream) {
    if (!db || (n_jobs && !jobs)) return set_error(PM_ERR_ARG, "pm_hintgen_dev: null pointer");
    int rc = ensure_device(db->device);
    if (rc) return rc;
    return hintgen_enqueue(db, jobs, n_jobs, stream ? (cudaStream_t)stream : db->stream);
}

// Host-buffer variant: tags/skip uploaded, parities downloaded.  Jobs are issued in launch groups so that the D2H of one
// group's parities overlaps the next group's kernel.  The shared last round makes a launch cost exactly its share of
// the work, so the groups can be as small as one job: the first copy starts after 1/n_jobs of the kernel time and the
// PCIe link stays busy from then on (350 MB of parities at ~56 GB/s is the longer leg of the MS-MARCO call).
PM_EXPORT int pm_hintgen(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs) {
    if (!db || (n_jobs && !jobs)) return set_error(PM_ERR_ARG, "pm_hintgen: null pointer");
    int rc = ensure_device(db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    const uint64_t E = db->entry_u64;
    uint64_t total_hints = 0, explicit_tags = 0, explicit_skip = 0;
    for (uint64_t a = 0; a < n_jobs; a++) {
        total_hints += jobs[a].n_hints;
        if (jobs[a].tags) explicit_tags += jobs[a].n_hints;
        if (jobs[a].skip_chunk) explicit_skip += jobs[a].n_hints;
    }
    if (total_hints == 0) return PM_OK;
    void *d_out = nullptr, *d_tags = nullptr, *d_skip = nullptr;
    if ((rc = scratch(db, 0, total_hints * E * 8, &d_out))) return rc;
    if (explicit_tags && (rc = scratch(db, 1, explicit_tags * 8, &d_tags))) return rc;
    if (explicit_skip && (rc = scratch(db, 2, explicit_skip * 4, &d_skip))) return rc;

    std::vector<pm_hint_job> dj(jobs, jobs + n_jobs);
    uint64_t ho = 0, to = 0, so = 0;
    for (uint64_t a = 0; a < n_jobs; a++) {
        dj[a].offsets_out = nullptr;   // a device-side output: not part of the host-buffer call
        dj[a].parity_out = (uint64_t *)d_out + ho * E;
        ho += jobs[a].n_hints;
        if (jobs[a].n_hints && !jobs[a].parity_out) return set_error(PM_ERR_ARG, "pm_hintgen: parity_out is null");
        if (jobs[a].tags) {
            dj[a].tags = (uint64_t *)d_tags + to;
            PM_CUDA(cudaMemcpyAsync((void *)dj[a].tags, jobs[a].tags, jobs[a].n_hints * 8, cudaMemcpyHostToDevice, db->stream));
            to += jobs[a].n_hints;
        }
        if (jobs[a].skip_chunk) {
            dj[a].skip_chunk = (int32_t *)d_skip + so;
            PM_CUDA(cudaMemcpyAsync((void *)dj[a].skip_chunk, jobs[a].skip_chunk, jobs[a].n_hints * 4, cudaMemcpyHostToDevice, db->stream));
            so += jobs[a].n_hints;
        }
    }
    // launch groups: one per job while a job is at least a machine-wide round of work, else four at most
    uint64_t groups = std::min<uint64_t>(n_jobs, 4);
    if (total_hints / n_jobs >= (uint64_t)db->sm_count * 64) groups = std::min<uint64_t>(n_jobs, 16);
    const int force_groups = tune(T_HG_D2H_GROUPS);
    if (force_groups >= 1) groups = std::min<uint64_t>(n_jobs, (uint64_t)force_groups);
    for (uint64_t gi = 0; gi < groups; gi++) {
        const uint64_t a0 = n_jobs * gi / groups, a1 = n_jobs * (gi + 1) / groups;
        if ((rc = hintgen_enqueue(db, dj.data() + a0, a1 - a0, db->stream))) return rc;
        PM_CUDA(cudaEventRecord(db->ev[gi & 3], db->stream));
        PM_CUDA(cudaStreamWaitEvent(db->copy_stream, db->ev[gi & 3], 0));
        for (uint64_t a = a0; a < a1; a++)
            if (jobs[a].n_hints)
                PM_CUDA(cudaMemcpyAsync(jobs[a].parity_out, dj[a].parity_out, jobs[a].n_hints * E * 8, cudaMemcpyDeviceToHost, db->copy_stream));
    }
    PM_CUDA(cudaStreamSynchronize(db->copy_stream));
    PM_CUDA(cudaStreamSynchronize(db->stream));
    return PM_OK;
}
