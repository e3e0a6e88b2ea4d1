// shard_multi.cu -- PatternModel::train on several GPUs of one node from ONE process, behind the C ABI (colibri_b200_train_multi).
//
// The torchrun-launched runs (colibri-core_b200/multigpu.py, bench.py --gpus N) drive the shard phases of shard.cu / shard_p2p.cu from one
// process per GPU.  A C++ caller -- the colibri-patternmodeller CLI with `-d 0-7`, or a program linking the host mirror -- has one process:
// here one host thread per device runs the same phases, the "symmetric" receive buffers are plain cudaMalloc blocks made visible to the
// other devices with cudaDeviceEnablePeerAccess (one address space: a pointer is a pointer), and a host barrier stands where the
// torchrun variant has a device-side barrier.  The split / reply kernels store straight into the owners' / senders' buffers over NVLink
// exactly as there.  Level 1 and the header numbers are reduced through host memory (a few hundred KB); the dense square of level 2 is
// summed by a kernel that reads the peers' squares.  A rank that fails raises a shared flag that every barrier checks, so the others
// leave instead of waiting forever.
// Not here (refused): exhaustive skipgrams (their exchange is an NCCL all-to-all in the torchrun variant) and indexed models.  MINLENGTH > 1 is
// handled by the drop rules of shard_finish (shard.cu).
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>

#include "shard.h"

using namespace colibri;

namespace {

struct HostBarrier {
    std::mutex              mu;
    std::condition_variable cv;
    int                     n, waiting = 0;
    uint64_t                generation = 0;
    bool                    aborted = false;
    explicit HostBarrier(int parties) : n(parties) {}
    // returns false when some rank has failed: the caller leaves
    bool wait() {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return false;
        const uint64_t gen = generation;
        if (++waiting == n) {
            waiting = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        cv.wait(lk, [&] { return generation != gen || aborted; });
        return !aborted;
    }
    void abort() {
        std::lock_guard<std::mutex> lk(mu);
        aborted = true;
        cv.notify_all();
    }
};

__global__ void __launch_bounds__(256) sum_squares_kernel(uint32_t* const* __restrict__ squares, uint32_t world, uint64_t n, uint32_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t s = 0;
        for (uint32_t r = 0; r < world; ++r) s += squares[r][i];
        out[i] = s;
    }
}

struct Rank {
    int                  dev = 0;
    colibri_b200_corpus* corpus = nullptr;
    colibri_b200_shard*  shard = nullptr;
    void *keys_rx = nullptr, *reply_rx = nullptr, *surv_rx = nullptr, *hdr = nullptr, *dense_local = nullptr, *dense_global = nullptr, *d_squares = nullptr;
    void*                d_counts = nullptr;
    std::vector<uint32_t> h_counts;
    uint64_t             tokens = 0, maxclass = 0;
    int                  rc = 0;
    std::string          err;
};

}  // namespace

extern "C" int colibri_b200_train_multi(const uint8_t* host_body, size_t nbytes, const colibri_b200_options* opt, const int* devices, int ndev, colibri_b200_model** out) {
    if (!opt || !devices || !out || ndev < 1 || ndev > 64) return set_err(COLIBRI_E_INVALID, "train_multi: bad arguments (1 .. 64 devices)");
    for (int r = 0; r < ndev; ++r) out[r] = nullptr;
    if (!host_body || nbytes == 0) return set_err(COLIBRI_E_FORMAT, "Attempting to read pattern from file, but file is empty?");
    colibri_b200_options o = *opt;
    TRY(check_options(o));
    if (o.DOSKIPGRAMS_EXHAUSTIVE || o.DOSKIPGRAMS) return set_err(COLIBRI_E_UNSUPPORTED, "skipgrams are not on the in-process multi-GPU path (use the torchrun variant, colibri-core_b200/multigpu.py)");
    const int have = colibri_b200_device_count();
    for (int r = 0; r < ndev; ++r) {
        if (devices[r] < 0 || devices[r] >= have) return set_err(COLIBRI_E_INVALID, "device %d out of range (have %d)", devices[r], have);
        for (int q = 0; q < r; ++q)
            if (devices[q] == devices[r]) return set_err(COLIBRI_E_INVALID, "device %d listed twice", devices[r]);
    }
    const uint32_t G = (uint32_t)ndev;

    // ---- cut the corpus at sentence boundaries (a delimiter is a 0x00 byte whose predecessor is below 128)
    std::vector<size_t> cuts(1, 0);
    for (uint32_t r = 1; r < G; ++r) {
        size_t j = std::max(nbytes * r / G, cuts.back());
        while (j < nbytes && !(host_body[j] == 0 && (j == 0 || host_body[j - 1] < 128))) ++j;
        cuts.push_back(std::min(j + 1, nbytes));
    }
    cuts.push_back(nbytes);
    size_t max_shard = 0;
    for (uint32_t r = 0; r < G; ++r) {
        if (cuts[r + 1] == cuts[r]) return set_err(COLIBRI_E_INVALID, "the corpus has too few sentences for %u devices", G);
        max_shard = std::max(max_shard, cuts[r + 1] - cuts[r]);
    }
    // a shard has at most one position per byte (+ the virtual delimiter); hash partitioning spreads a level's windows evenly over the owners
    const uint64_t slot_cap = (uint64_t)(1.3 * (double)(max_shard + 1) / G) + 65536;
    const uint64_t surv_cap = slot_cap / 2 + 4096;
    const Tuning   tune     = Tuning::from_env();

    std::vector<Rank> ranks(G);
    HostBarrier       bar((int)G);
    // shared between the ranks (written before a barrier, read after it)
    std::vector<uint64_t> g_tokens(G, 0), g_maxclass(G, 0);
    std::vector<uint32_t> g_counts;  // the summed class histogram
    std::mutex            g_mu;

    // ---- peer access (both directions, every pair)
    for (uint32_t a = 0; a < G; ++a) {
        CUDA_TRY(cudaSetDevice(devices[a]));
        for (uint32_t b = 0; b < G; ++b) {
            if (a == b) continue;
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
            if (!can) return set_err(COLIBRI_E_UNSUPPORTED, "device %d cannot access device %d: no peer path", devices[a], devices[b]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_err(COLIBRI_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[a], devices[b], cudaGetErrorString(e));
            cudaGetLastError();
        }
    }

    auto fail = [&](Rank& rk, int rc) {
        rk.rc  = rc ? rc : COLIBRI_E_CUDA;
        rk.err = colibri_b200_last_error();
        bar.abort();
    };

    // ---- set-up per rank (sequential: allocations, staging, tokenising)
    for (uint32_t r = 0; r < G; ++r) {
        Rank& rk = ranks[r];
        rk.dev   = devices[r];
        colibri_b200_options ro = o;
        ro.device               = rk.dev;
        TRY(colibri_b200_corpus_stage(host_body + cuts[r], cuts[r + 1] - cuts[r], rk.dev, &rk.corpus));
        TRY(colibri_b200_shard_begin(rk.corpus, &ro, (int)r, (int)G, &rk.shard));
        CUDA_TRY(cudaSetDevice(rk.dev));
        CUDA_TRY(cudaMalloc(&rk.keys_rx, G * slot_cap * 8));
        CUDA_TRY(cudaMalloc(&rk.reply_rx, G * slot_cap * 4));
        CUDA_TRY(cudaMalloc(&rk.surv_rx, G * surv_cap * 8));
        CUDA_TRY(cudaMalloc(&rk.hdr, (6 * G + 64) * 8));
        CUDA_TRY(cudaMemset(rk.hdr, 0, (6 * G + 64) * 8));
        uint64_t info[4];
        TRY(colibri_b200_shard_info(rk.shard, info));
        rk.tokens   = info[0];
        rk.maxclass = info[1];
    }
    uint64_t global_tokens = 0, maxclass = 0;
    for (auto& rk : ranks) {
        global_tokens += rk.tokens;
        maxclass = std::max(maxclass, rk.maxclass);
    }
    const uint32_t nclasses = (uint32_t)maxclass + 1;
    uint32_t       dense    = 0;
    if (tune.dense_dim >= 2 && global_tokens / G >= tune.dense_min) dense = std::min<uint32_t>(std::min<uint32_t>(tune.dense_dim, nclasses), 16384);
    const uint64_t dense_cells = (uint64_t)dense * dense;
    {
        std::vector<uint64_t> pk(G), pr(G), ps(G), ph(G);
        for (uint32_t r = 0; r < G; ++r) {
            pk[r] = (uint64_t)(uintptr_t)ranks[r].keys_rx;
            pr[r] = (uint64_t)(uintptr_t)ranks[r].reply_rx;
            ps[r] = (uint64_t)(uintptr_t)ranks[r].surv_rx;
            ph[r] = (uint64_t)(uintptr_t)ranks[r].hdr;
        }
        for (uint32_t r = 0; r < G; ++r) {
            Rank& rk = ranks[r];
            TRY(colibri_b200_shard_set_peers(rk.shard, pk.data(), pr.data(), ps.data(), ph.data(), slot_cap, surv_cap));
            CUDA_TRY(cudaSetDevice(rk.dev));
            CUDA_TRY(cudaMalloc(&rk.d_counts, (size_t)nclasses * 4));
            rk.h_counts.resize(nclasses);
            if (dense) {
                CUDA_TRY(cudaMalloc(&rk.dense_local, dense_cells * 4));
                CUDA_TRY(cudaMalloc(&rk.dense_global, dense_cells * 4));
                CUDA_TRY(cudaMalloc(&rk.d_squares, G * sizeof(void*)));
                CUDA_TRY(cudaMemset(rk.dense_local, 0, dense_cells * 4));
            }
        }
        if (dense) {
            std::vector<void*> sq(G);
            for (uint32_t r = 0; r < G; ++r) sq[r] = ranks[r].dense_local;
            for (uint32_t r = 0; r < G; ++r) {
                CUDA_TRY(cudaSetDevice(ranks[r].dev));
                CUDA_TRY(cudaMemcpy(ranks[r].d_squares, sq.data(), G * sizeof(void*), cudaMemcpyHostToDevice));
                TRY(colibri_b200_shard_set_dense(ranks[r].shard, ranks[r].dense_local, dense));
            }
        }
    }

    // ---- the levels, one thread per rank
    struct Result {
        std::vector<uint64_t> passes;  // 4 per pass
        uint64_t              types = 0;
        int                   maxn = 0, minn = 999;
    };
    std::vector<Result> results(G);
    auto body = [&](uint32_t r) {
        Rank&   rk = ranks[r];
        Result& rs = results[r];
        cudaSetDevice(rk.dev);
        colibri_b200_shard* sh = rk.shard;
        int rc;
        // level 1: class histograms summed through host memory
        if ((rc = colibri_b200_shard_unigram_counts(sh, nclasses, rk.d_counts)) != 0) return fail(rk, rc);
        if (cudaMemcpy(rk.h_counts.data(), rk.d_counts, (size_t)nclasses * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(rk, set_err(COLIBRI_E_CUDA, "D2H of the class histogram failed"));
        if (!bar.wait()) return;
        if (r == 0) {
            g_counts.assign(nclasses, 0);
            for (auto& other : ranks)
                for (uint32_t c = 0; c < nclasses; ++c) g_counts[c] += other.h_counts[c];
        }
        if (!bar.wait()) return;
        if (cudaMemcpy(rk.d_counts, g_counts.data(), (size_t)nclasses * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail(rk, set_err(COLIBRI_E_CUDA, "H2D of the class histogram failed"));
        uint64_t st[3];
        if ((rc = colibri_b200_shard_unigram_finish(sh, rk.d_counts, global_tokens, st)) != 0) return fail(rk, rc);
        const uint64_t found1 = st[0], kept1 = st[1];
        rs.types = found1;
        if (found1) {
            rs.passes.insert(rs.passes.end(), {1, found1, 0, found1 - kept1});
            rs.maxn = rs.minn = 1;
        }
        uint64_t prev_kept = kept1;
        for (int n = 2; found1 && n <= o.MAXLENGTH && prev_kept > 0; ++n) {
            uint64_t windows = 0;
            if ((rc = colibri_b200_shard_p2p_split(sh, n, &windows)) != 0) return fail(rk, rc);
            if (n == 2 && dense) {
                // every rank sums all squares into its own copy (the peers' squares are read over NVLink), then reads the verdicts from that copy
                if (!bar.wait()) return;
                sum_squares_kernel<<<296, 256, 0, sh->s>>>((uint32_t* const*)rk.d_squares, G, dense_cells, (uint32_t*)rk.dense_global);
                if (cudaStreamSynchronize(sh->s) != cudaSuccess) return fail(rk, set_err(COLIBRI_E_CUDA, "summing the dense squares failed: %s", cudaGetErrorString(cudaGetLastError())));
                sh->dense_cnt = (uint32_t*)rk.dense_global;
            }
            if (!bar.wait()) return;
            uint64_t ost[3];
            if ((rc = colibri_b200_shard_p2p_owner(sh, ost)) != 0) return fail(rk, rc);
            if (!bar.wait()) return;
            uint64_t gst[3], valid = 0;
            if ((rc = colibri_b200_shard_p2p_finish(sh, gst, &valid)) != 0) return fail(rk, rc);
            if (!bar.wait()) return;  // the receive buffers are free for the next level
            if (gst[0] == 0) break;   // "None found" (reference include/patternmodel.h:1189-1194)
            rs.passes.insert(rs.passes.end(), {(uint64_t)n, gst[0], 0, gst[0] - gst[1]});
            rs.maxn = std::max(rs.maxn, n);
            rs.minn = std::min(rs.minn, n);
            prev_kept = gst[1];
        }
        if (o.MINTOKENS == 1 && !rs.passes.empty()) {  // the reference reports one pass when every length is extracted in a single scan
            uint64_t f = 0, k = 0;
            for (size_t p = 0; p < rs.passes.size(); p += 4) {
                f += rs.passes[p + 1];
                k += rs.passes[p + 3];
            }
            rs.passes = {1, f, 0, k};
        }
        if ((rc = colibri_b200_shard_finish(sh, rs.passes.data(), (int)(rs.passes.size() / 4), rs.types, rs.maxn, rs.minn, &out[r])) != 0) return fail(rk, rc);
    };
    {
        std::vector<std::thread> threads;
        for (uint32_t r = 0; r < G; ++r) threads.emplace_back(body, r);
        for (auto& t : threads) t.join();
    }

    // ---- tear-down
    int         rc = 0;
    std::string err;
    for (uint32_t r = 0; r < G; ++r) {
        Rank& rk = ranks[r];
        if (rk.rc && !rc) {
            rc  = rk.rc;
            err = rk.err;
        }
        cudaSetDevice(rk.dev);
        colibri_b200_shard_free(rk.shard);
        colibri_b200_corpus_free(rk.corpus);
        for (void* p : {rk.keys_rx, rk.reply_rx, rk.surv_rx, rk.hdr, rk.dense_local, rk.dense_global, rk.d_squares, rk.d_counts}) cudaFree(p);
    }
    if (rc) {
        for (uint32_t r = 0; r < G; ++r) {
            colibri_b200_model_free(out[r]);
            out[r] = nullptr;
        }
        return set_err(rc, "train_multi: %s", err.c_str());
    }
    return 0;
}
