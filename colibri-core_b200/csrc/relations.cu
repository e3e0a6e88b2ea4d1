// relations.cu -- queries on the reverse index of a model over a corpus (colibri_b200_rindex, built in engine.cu: Trainer::build_rindex).
//
//   rindex_query     IndexedPatternModel::getreverseindex (reference include/patternmodel.h:1746-1824), batched: the model's n-grams that start at
//                    each of nq (sentence, token) references -- a gather from the match arrays
//   rindex_cooc      getrightcooc / getleftcooc (:3460-3493, :3502-3531) for ALL patterns at once.  The reference reaches the neighbouring
//                    positions through getreverseindex_right / _left (:1867-1878, :1885-1892), which look every neighbouring position ref2 up at
//                    the ORIGINAL reference ref; so what it computes -- and what is reproduced here, bit for bit -- is: a relation (P, Q) exists for
//                    two model patterns that start at the SAME corpus position (s, t), and each such position adds
//                        right:  max(0, sl - 1 - (t + |P|))   (the positions i of the sentence with i > t + |P|)
//                        left:   max(0, t - |Q|)              (the positions i < t with i + |Q| < t)
//                    to joint(P, Q).  One thread per corpus position, one 16-byte-slot table keyed by (index of P, index of Q), 64-bit sums.
//   rindex_cooc_of   getcooc (:3543-3576) of ONE pattern P: every occurrence of P against every model n-gram that starts anywhere in the same
//                    sentence (getreverseindex_bysentence :1850-1862) and neither overlaps nor touches it: t2 + |Q| < t or t2 > t + |P|.  One thread
//                    per occurrence of P (a scan of P's match column finds them), a dense counter per model pattern, compacted afterwards.
// computenpmi (:3671-3691) and computeflexgrams_fromcooc (:3751-3774) are floating point / string work on these integers and live on the host
// side (host/patternmodel.h, colibri-core_b200/pybinding.py), with the reference's own expression for npmi() (:3582-3585).
#include <memory>
#include <mutex>
#include <unordered_map>

#include "device_utils.cuh"
#include "engine_common.h"
#include "relations.h"

using namespace colibri;

namespace {

__global__ void __launch_bounds__(256) rindex_query_kernel(const uint32_t* __restrict__ sentence, const uint16_t* __restrict__ token, uint64_t nq,
                                                           const uint32_t* __restrict__ sent_start, uint64_t nsentences, const uint32_t* const* __restrict__ match,
                                                           uint32_t nlen, uint32_t* __restrict__ out) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint32_t s = sentence[q], t = token[q];
    bool           ok = s >= 1 && s <= nsentences;
    uint32_t       p = 0;
    if (ok) {
        const uint32_t first = sent_start[s - 1], sl = sent_start[s] - 1 - first;  // the sentence's delimiter is not a token of it
        ok                   = t < sl;
        p                    = first + t;
    }
    for (uint32_t k = 0; k < nlen; ++k) out[q * nlen + k] = ok ? __ldg(match[k] + p) : 0u;
}

struct alignas(16) CoocSlot {
    unsigned long long key;  // (index of P + 1) << 32 | (index of Q + 1); 0 = empty
    unsigned long long sum;
};

__device__ __forceinline__ void cooc_add(CoocSlot* __restrict__ table, uint64_t cap, unsigned long long key, unsigned long long w, bool& full) {
    uint64_t slot = fast_range(table_hash_u64(key), cap);
    for (uint64_t step = 0; step < cap && step < 8192; ++step) {
        unsigned long long cur = table[slot].key;
        if (cur == 0) {
            cur = atomicCAS(&table[slot].key, 0ull, key);
            if (cur == 0) cur = key;
        }
        if (cur == key) {
            atomicAdd(&table[slot].sum, w);
            return;
        }
        slot = slot + 1 == cap ? 0 : slot + 1;
    }
    full = true;
}

constexpr int kMaxLens = 32;

__global__ void __launch_bounds__(256) rindex_cooc_kernel(const uint32_t* const* __restrict__ match, const uint32_t* __restrict__ lengths, uint32_t nlen, uint64_t npos,
                                                          const uint64_t* __restrict__ sent_before, const uint32_t* __restrict__ sent_start, int left, CoocSlot* __restrict__ table,
                                                          uint64_t cap, DeviceStats* __restrict__ st) {
    bool full = false;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t ids[kMaxLens];
        uint32_t any = 0;
        for (uint32_t k = 0; k < nlen; ++k) {
            ids[k] = __ldg(match[k] + p);
            any |= ids[k];
        }
        if (!any) continue;
        const uint64_t s     = sent_before[p];  // 0-based sentence of p (delimiters before it)
        const uint32_t first = sent_start[s], t = (uint32_t)(p - first), sl = sent_start[s + 1] - 1 - first;
        for (uint32_t i = 0; i < nlen; ++i) {
            if (!ids[i]) continue;
            for (uint32_t j = 0; j < nlen; ++j) {
                if (!ids[j]) continue;
                // P = the pattern of lengths[i], Q = its neighbour of lengths[j], both at (s, t)
                const int64_t w = left ? (int64_t)t - (int64_t)lengths[j] : (int64_t)sl - 1 - ((int64_t)t + (int64_t)lengths[i]);
                if (w > 0) cooc_add(table, cap, ((unsigned long long)ids[i] << 32) | ids[j], (unsigned long long)w, full);
            }
        }
    }
    if (full) atomicOr(&st->errflags, kErrTableFull);
}

__global__ void __launch_bounds__(256) cooc_collect_kernel(const CoocSlot* __restrict__ table, uint64_t cap, uint32_t* __restrict__ idx_p, uint32_t* __restrict__ idx_q,
                                                           unsigned long long* __restrict__ joint, uint64_t out_cap, DeviceStats* __restrict__ st) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (cap + 31) / 32 * 32; i += (uint64_t)gridDim.x * blockDim.x) {
        const bool     used = i < cap && table[i].key != 0;
        const uint64_t o    = warp_aggregated_inc(&st->cursor, used);
        if (used && idx_p != nullptr && o < out_cap) {
            idx_p[o] = (uint32_t)(table[i].key >> 32) - 1;
            idx_q[o] = (uint32_t)table[i].key - 1;
            joint[o] = table[i].sum;
        }
    }
}

// the counters of the last colibri_b200_rindex_cooc_of query of an index (see relations.h)
struct CoocOf {
    DevBuf<unsigned long long> counts;       // co-occurrence count per model pattern of ...
    uint64_t                   pattern = 0;  // ... this pattern (index + 1; 0 = none)
};
std::mutex                                                            g_cooc_of_mu;
std::unordered_map<const colibri_b200_rindex*, std::unique_ptr<CoocOf>> g_cooc_of;

__global__ void __launch_bounds__(256) rindex_cooc_of_kernel(const uint32_t* const* __restrict__ match, const uint32_t* __restrict__ lengths, uint32_t nlen, uint64_t npos,
                                                             const uint64_t* __restrict__ sent_before, const uint32_t* __restrict__ sent_start, uint32_t id_p /* index + 1 */,
                                                             uint32_t col_p, unsigned long long* __restrict__ counts) {
    const uint32_t n_p = lengths[col_p];
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (uint64_t)gridDim.x * blockDim.x) {
        if (__ldg(match[col_p] + p) != id_p) continue;
        const uint64_t s     = sent_before[p];
        const uint32_t first = sent_start[s], t = (uint32_t)(p - first), sl = sent_start[s + 1] - 1 - first;
        for (uint32_t t2 = 0; t2 < sl; ++t2) {
            if (t2 <= t + n_p && t2 + 1 >= t) continue;  // even a unigram here would touch or overlap P
            for (uint32_t k = 0; k < nlen; ++k) {
                const uint32_t q = __ldg(match[k] + first + t2);
                if (q && (t2 + lengths[k] < t || t2 > t + n_p)) atomicAdd(counts + (q - 1), 1ull);
            }
        }
    }
}

__global__ void __launch_bounds__(256) cooc_of_collect_kernel(const unsigned long long* __restrict__ counts, uint64_t npatterns, uint32_t* __restrict__ idx_q,
                                                              unsigned long long* __restrict__ out, uint64_t out_cap, DeviceStats* __restrict__ st) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (npatterns + 31) / 32 * 32; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long c    = i < npatterns ? counts[i] : 0ull;
        const uint64_t           o    = warp_aggregated_inc(&st->cursor, c != 0);
        if (c != 0 && idx_q != nullptr && o < out_cap) {
            idx_q[o] = (uint32_t)i;
            out[o]   = c;
        }
    }
}

}  // namespace

void colibri::rindex_cooc_forget(const colibri_b200_rindex* r) {
    std::lock_guard<std::mutex> g(g_cooc_of_mu);
    g_cooc_of.erase(r);
}

extern "C" int colibri_b200_rindex_lengths(const colibri_b200_rindex* r, uint32_t* lengths, uint32_t cap, uint32_t* n) {
    if (!r || !n) return set_err(COLIBRI_E_INVALID, "NULL argument");
    *n = (uint32_t)r->lengths.size();
    if (lengths)
        for (uint32_t k = 0; k < *n && k < cap; ++k) lengths[k] = (uint32_t)r->lengths[k];
    return 0;
}

extern "C" int colibri_b200_rindex_sentence_starts(const colibri_b200_rindex* r, uint32_t* out, uint64_t cap) {
    if (!r || !out) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (cap < r->nsentences + 1) return set_err(COLIBRI_E_INVALID, "buffer too small: %llu entries needed", (unsigned long long)(r->nsentences + 1));
    CUDA_TRY(cudaSetDevice(r->device));
    CUDA_TRY(cudaMemcpyAsync(out, r->sent_start.p, (r->nsentences + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, r->stream));
    CUDA_TRY(cudaStreamSynchronize(r->stream));
    return 0;
}

extern "C" int colibri_b200_rindex_query(colibri_b200_rindex* r, const uint32_t* sentence, const uint16_t* token, uint64_t nq, uint32_t* out) {
    if (!r || (nq && (!sentence || !token || !out))) return set_err(COLIBRI_E_INVALID, "NULL argument");
    const uint32_t nlen = (uint32_t)r->lengths.size();
    if (nq == 0 || nlen == 0) return 0;
    CUDA_TRY(cudaSetDevice(r->device));
    cudaStream_t     s = r->stream;
    DevBuf<uint32_t> d_s, d_out;
    DevBuf<uint16_t> d_t;
    TRY(d_s.alloc(r->device, nq));
    TRY(d_t.alloc(r->device, nq));
    TRY(d_out.alloc(r->device, nq * nlen));
    CUDA_TRY(cudaMemcpyAsync(d_s.p, sentence, nq * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_t.p, token, nq * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
    rindex_query_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(d_s.p, d_t.p, nq, r->sent_start.p, r->nsentences, r->d_match_ptrs.p, nlen, d_out.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, d_out.p, nq * nlen * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int colibri_b200_rindex_cooc(colibri_b200_rindex* r, int direction, uint32_t* idx_p, uint32_t* idx_q, uint64_t* joint, uint64_t cap_out, uint64_t* nrel) {
    if (!r || !nrel) return set_err(COLIBRI_E_INVALID, "NULL argument");
    if (direction != 0 && direction != 1) return set_err(COLIBRI_E_INVALID, "direction %d (0 = right, 1 = left)", direction);
    *nrel = 0;
    const uint32_t nlen = (uint32_t)r->lengths.size();
    if (nlen == 0 || r->npos == 0) return 0;
    if (nlen > kMaxLens) return set_err(COLIBRI_E_UNSUPPORTED, "co-occurrence over %u pattern lengths (at most %d)", nlen, kMaxLens);
    CUDA_TRY(cudaSetDevice(r->device));
    cudaStream_t s = r->stream;
    int          sms = 148;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, r->device));
    // two patterns that start at one position are prefix and extension of each other: at most (2 * (lengths - 1) + 1) relations per pattern
    uint64_t cap = std::max<uint64_t>(1024, (uint64_t)(r->model->npatterns * (2ull * nlen - 1)) * 3 / 2 + 64);
    DevBuf<DeviceStats> d_stats;
    DeviceStats         h;
    TRY(d_stats.alloc(r->device, 1));
    DevBuf<CoocSlot> table;
    for (;;) {
        TRY(table.alloc(r->device, cap));
        CUDA_TRY(cudaMemsetAsync(table.p, 0, cap * sizeof(CoocSlot), s));
        CUDA_TRY(cudaMemsetAsync(d_stats.p, 0, sizeof(DeviceStats), s));
        const unsigned grid = (unsigned)std::min<uint64_t>((r->npos + 255) / 256, (uint64_t)sms * 16);
        rindex_cooc_kernel<<<grid, 256, 0, s>>>(r->d_match_ptrs.p, r->d_lengths.p, nlen, r->npos, r->sent_before.p, r->sent_start.p, direction, table.p, cap, d_stats.p);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&h, d_stats.p, sizeof h, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (!(h.errflags & kErrTableFull)) break;
        if (cap > (1ull << 36)) return set_err(COLIBRI_E_CAPACITY, "co-occurrence table overflow");
        cap *= 2;
    }
    DevBuf<uint32_t>           d_p, d_q;
    DevBuf<unsigned long long> d_j;
    const bool want = idx_p && idx_q && joint && cap_out;
    if (want) {
        TRY(d_p.alloc(r->device, cap_out));
        TRY(d_q.alloc(r->device, cap_out));
        TRY(d_j.alloc(r->device, cap_out));
    }
    CUDA_TRY(cudaMemsetAsync(&d_stats.p->cursor, 0, sizeof(unsigned long long), s));
    const unsigned grid = (unsigned)std::min<uint64_t>((cap + 255) / 256, (uint64_t)sms * 16);
    cooc_collect_kernel<<<grid, 256, 0, s>>>(table.p, cap, want ? d_p.p : nullptr, d_q.p, d_j.p, cap_out, d_stats.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(&h, d_stats.p, sizeof h, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *nrel = h.cursor;
    if (want) {
        if (h.cursor > cap_out) return set_err(COLIBRI_E_CAPACITY, "%llu relations; the buffers hold %llu", (unsigned long long)h.cursor, (unsigned long long)cap_out);
        CUDA_TRY(cudaMemcpyAsync(idx_p, d_p.p, h.cursor * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(idx_q, d_q.p, h.cursor * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(joint, d_j.p, h.cursor * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    return 0;
}

extern "C" int colibri_b200_rindex_cooc_of(colibri_b200_rindex* r, uint64_t pattern, uint32_t* idx_q, uint64_t* count, uint64_t cap_out, uint64_t* nrel) {
    if (!r || !nrel) return set_err(COLIBRI_E_INVALID, "NULL argument");
    *nrel = 0;
    colibri_b200_model* m = r->model;
    if (pattern >= m->npatterns) return set_err(COLIBRI_E_INVALID, "pattern index %llu of %llu", (unsigned long long)pattern, (unsigned long long)m->npatterns);
    const uint32_t nlen = (uint32_t)r->lengths.size();
    CUDA_TRY(cudaSetDevice(r->device));
    cudaStream_t s = r->stream;
    int          sms = 148;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, r->device));
    CoocOf* cache;
    {
        std::lock_guard<std::mutex> g(g_cooc_of_mu);
        std::unique_ptr<CoocOf>&    slot = g_cooc_of[r];
        if (!slot) slot.reset(new CoocOf());
        cache = slot.get();  // (an index handle is used by one host thread at a time, include/colibri_b200.h)
    }
    if (cache->pattern != pattern + 1) {  // (a caller asks twice: for the number of relations, then for the relations)
        cache->pattern = 0;
        TRY(ensure_meta(m, nullptr));
        uint16_t n_p = 0;
        uint8_t  cat = 0;
        CUDA_TRY(cudaMemcpyAsync(&n_p, m->d_pn.p + pattern, sizeof n_p, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(&cat, m->d_pcat.p + pattern, sizeof cat, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (cache->counts.n < m->npatterns) TRY(cache->counts.alloc(r->device, m->npatterns));
        CUDA_TRY(cudaMemsetAsync(cache->counts.p, 0, m->npatterns * sizeof(unsigned long long), s));
        uint32_t col = nlen;
        for (uint32_t k = 0; k < nlen; ++k)
            if ((uint32_t)r->lengths[k] == n_p) col = k;
        if (cat == 0 && col < nlen && r->npos) {  // (only n-grams are matched against the corpus: a skipgram or flexgram has no occurrence here)
            const unsigned grid = (unsigned)std::min<uint64_t>((r->npos + 255) / 256, (uint64_t)sms * 16);
            rindex_cooc_of_kernel<<<grid, 256, 0, s>>>(r->d_match_ptrs.p, r->d_lengths.p, nlen, r->npos, r->sent_before.p, r->sent_start.p, (uint32_t)pattern + 1, col,
                                                       cache->counts.p);
            CUDA_TRY(cudaGetLastError());
        }
        cache->pattern = pattern + 1;
    }
    DevBuf<DeviceStats>        d_stats;
    DevBuf<uint32_t>           d_q;
    DevBuf<unsigned long long> d_c;
    DeviceStats                h;
    TRY(d_stats.alloc(r->device, 1));
    const bool want = idx_q && count && cap_out;
    if (want) {
        TRY(d_q.alloc(r->device, cap_out));
        TRY(d_c.alloc(r->device, cap_out));
    }
    CUDA_TRY(cudaMemsetAsync(d_stats.p, 0, sizeof(DeviceStats), s));
    const unsigned grid = (unsigned)std::min<uint64_t>((m->npatterns + 255) / 256, (uint64_t)sms * 16);
    cooc_of_collect_kernel<<<grid, 256, 0, s>>>(cache->counts.p, m->npatterns, want ? d_q.p : nullptr, d_c.p, cap_out, d_stats.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(&h, d_stats.p, sizeof h, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *nrel = h.cursor;
    if (want) {
        if (h.cursor > cap_out) return set_err(COLIBRI_E_CAPACITY, "%llu relations; the buffers hold %llu", (unsigned long long)h.cursor, (unsigned long long)cap_out);
        CUDA_TRY(cudaMemcpyAsync(idx_q, d_q.p, h.cursor * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(count, d_c.p, h.cursor * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    return 0;
}
