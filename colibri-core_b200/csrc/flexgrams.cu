// flexgrams.cu -- flexgrams abstracted from the skipgrams of an indexed model (first piece).
//
// Reference: IndexedPatternModel::computeflexgrams_fromskipgrams (include/patternmodel.h:3724-3744): every skipgram is turned into its
// flexgram (Pattern::toflexgram, src/pattern.cpp:145-180: each run of gap tokens 0x03 becomes ONE dynamic gap 0x04) and all its
// occurrences are appended to that flexgram's list; the return value counts the flexgrams that were new.  The CLI does this for `-F S`
// after trainskipgrams (src/patternmodeller.cpp:330-337).  The reference inserts into the map it iterates over (undefined behaviour
// whenever that rehashes, see tests/golden/make_golden_flex.py); this is the clean iteration: every skipgram exactly once.
//
// On the device it is a group-by on variable-length keys: the flexgram bytes of all skipgrams are written into one blob, a pattern index
// (pattern_index.cu: SpookyV2 of the bytes, 32-byte inline-key slots) is built over it in the mode that resolves duplicates to the first
// claimant, the occurrence triples (group, sentence, token) are emitted per skipgram and sorted with the stable LSD radix sort of index.cu
// (token, then sentence, then group), so every flexgram's list comes out ascending, and the result is the old model with the flexgrams
// appended.  All streaming or random-sector HBM work; nothing here touches the corpus.
#include "device_utils.cuh"
#include "engine_common.h"

using namespace colibri;

namespace {

inline unsigned fx_div_up(uint64_t a, uint64_t b) {
    return (unsigned)((a + b - 1) / b);
}

__global__ void __launch_bounds__(256) flag_category_kernel(const uint8_t* __restrict__ pcat, uint64_t np, uint8_t cat, uint32_t* __restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) flags[i] = pcat[i] == cat ? 1u : 0u;
}
__global__ void __launch_bounds__(256) scatter_selected_kernel(const uint32_t* __restrict__ flags, const uint64_t* __restrict__ newpos, uint64_t n, uint32_t* __restrict__ list) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) list[newpos[i]] = (uint32_t)i;
}
// Pattern::toflexgram on the key of skipgram sel[j]: its length (out == NULL) or its bytes
__device__ __forceinline__ uint32_t toflexgram(const uint8_t* __restrict__ key, uint32_t len, uint8_t* __restrict__ out) {
    uint32_t j = 0;
    bool     skipgap = false, prevhigh = false;
    for (uint32_t i = 0; i < len; ++i) {
        const uint8_t c = key[i];
        if (!prevhigh && c == 3) {
            if (!skipgap) {
                if (out) out[j] = 4;
                ++j;
                skipgap = true;
            }
        } else {
            if (out) out[j] = c;
            ++j;
            skipgap = false;
        }
        prevhigh = c >= 128;
    }
    return j;
}
__global__ void __launch_bounds__(256) flex_lengths_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, const uint32_t* __restrict__ sel, uint64_t k,
                                                           const uint32_t* __restrict__ counts, uint32_t* __restrict__ lens, uint32_t* __restrict__ sel_counts) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const uint32_t i = sel[j];
    lens[j]       = toflexgram(keys + off[i], (uint32_t)(off[i + 1] - off[i]), nullptr);
    sel_counts[j] = counts[i];
}
__global__ void __launch_bounds__(256) flex_write_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ off, const uint32_t* __restrict__ sel, uint64_t k,
                                                         const uint64_t* __restrict__ foff, uint8_t* __restrict__ out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const uint32_t i = sel[j];
    toflexgram(keys + off[i], (uint32_t)(off[i + 1] - off[i]), out + foff[j]);
}
__global__ void __launch_bounds__(256) flag_representatives_kernel(const uint32_t* __restrict__ rep, uint64_t k, uint32_t* __restrict__ flags) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) flags[j] = rep[j] == (uint32_t)j ? 1u : 0u;
}
// group of skipgram j = number of its representative among the distinct flexgrams; the flexgram's count is the sum of its skipgrams' counts
__global__ void __launch_bounds__(256) flex_groups_kernel(const uint32_t* __restrict__ rep, const uint64_t* __restrict__ uniq_pos, const uint32_t* __restrict__ sel_counts, uint64_t k,
                                                          uint32_t* __restrict__ group, uint32_t* __restrict__ flex_counts) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const uint32_t g = (uint32_t)uniq_pos[rep[j]];
    group[j] = g;
    atomicAdd(&flex_counts[g], sel_counts[j]);
}
// one warp per skipgram: its occurrences, tagged with the flexgram they now also belong to
__global__ void __launch_bounds__(256) flex_emit_kernel(const uint32_t* __restrict__ sel, uint64_t k, const uint32_t* __restrict__ group, const uint64_t* __restrict__ emit_off,
                                                        const uint64_t* __restrict__ ref_off, const uint32_t* __restrict__ rs, const uint16_t* __restrict__ rt,
                                                        uint32_t* __restrict__ e_grp, uint32_t* __restrict__ e_sent, uint32_t* __restrict__ e_tok) {
    const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= k) return;
    const uint32_t i = sel[j], g = group[j];
    const uint64_t a = ref_off[i], l = ref_off[i + 1] - a, d = emit_off[j];
    for (uint64_t b = lane_id(); b < l; b += 32) {
        e_grp[d + b]  = g;
        e_sent[d + b] = rs[a + b];
        e_tok[d + b]  = rt[a + b];
    }
}
__global__ void __launch_bounds__(256) gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ perm, uint64_t n, uint32_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[perm[i]];
}
__global__ void __launch_bounds__(256) gather_refs_perm_kernel(const uint32_t* __restrict__ e_sent, const uint32_t* __restrict__ e_tok, const uint32_t* __restrict__ perm, uint64_t n,
                                                               uint32_t* __restrict__ rs_out, uint16_t* __restrict__ rt_out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        rs_out[i] = e_sent[perm[i]];
        rt_out[i] = (uint16_t)e_tok[perm[i]];
    }
}
__global__ void __launch_bounds__(256) max_u32_kernel(const uint32_t* __restrict__ v, uint64_t n, unsigned int* __restrict__ out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, v[i]);
    m = warp_reduce_max(m);
    if (lane_id() == 0 && m) atomicMax(out, m);
}
__global__ void __launch_bounds__(256) shift_u64_kernel(const uint64_t* __restrict__ src, uint64_t n, uint64_t add, uint64_t* __restrict__ dst) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}

// stable LSD radix sort of `perm` by field[perm], the bytes of the field that can be non-zero (maxval) only
int sort_perm_by(cudaStream_t s, const uint32_t* field, uint32_t maxval, uint64_t n, uint32_t*& perm, uint32_t*& perm_alt, uint32_t* keys, uint32_t* keys_alt, uint32_t* hist,
                 uint64_t* hist_off, uint64_t* stmp, uint64_t& launches) {
    gather_u32_kernel<<<fx_div_up(n, 256), 256, 0, s>>>(field, perm, n, keys);
    ++launches;
    uint32_t *kin = keys, *kout = keys_alt;
    for (int shift = 0; shift < 32 && (maxval >> shift) != 0; shift += 8) {
        launches += launch_radix_pass(s, kin, perm, n, shift, hist, hist_off, stmp, kout, perm_alt);
        std::swap(kin, kout);
        std::swap(perm, perm_alt);
    }
    return 0;
}

}  // namespace

extern "C" int colibri_b200_model_hasflexgrams(const colibri_b200_model* m) {
    return m ? m->hasflexgrams : 0;
}

extern "C" int colibri_b200_model_flexgrams_fromskipgrams(colibri_b200_model* m, uint64_t* found, colibri_b200_model** out) {
    if (!out) return set_err(COLIBRI_E_INVALID, "out is NULL");
    *out = nullptr;
    if (found) *found = 0;
    if (!m) return set_err(COLIBRI_E_INVALID, "model is NULL");
    if (m->model_type != COLIBRI_INDEXEDPATTERNMODEL || !m->d_ref_off.p)
        return set_err(COLIBRI_E_INVALID, "computeflexgrams_fromskipgrams needs an indexed model");  // the reference defines it for IndexedPatternModel only
    CUDA_TRY(cudaSetDevice(m->device));
    uint64_t launches = 0;
    TRY(ensure_meta(m, &launches));
    if (m->meta.hasflex) return set_err(COLIBRI_E_UNSUPPORTED, "the model already holds flexgrams; merging into existing flexgrams is not on the device path");
    if (m->meta.malformed) return set_err(COLIBRI_E_UNSUPPORTED, "%u pattern(s) are not indexable on the device", m->meta.malformed);
    colibri_b200_model* r = nullptr;
    TRY(new_model(m->device, COLIBRI_INDEXEDPATTERNMODEL, &r));
    const int      dev = m->device;
    cudaStream_t   s   = r->stream;
    const uint64_t np  = m->npatterns;
    auto body = [&]() -> int {
        // ---- the skipgrams
        DevBuf<uint32_t> flags, sel, lens, sel_counts, rep, uflags, uniq, group, flex_counts, flex_lens;
        DevBuf<uint64_t> pos, tmp, foff, upos, emit_off;
        DevBuf<uint8_t>  fkeys;
        uint64_t         K = 0;
        TRY(flags.alloc(dev, np + 1));
        TRY(pos.alloc(dev, np + 1));
        TRY(tmp.alloc(dev, np / 2048 + 4));
        if (np) {
            flag_category_kernel<<<fx_div_up(np, 256), 256, 0, s>>>(m->d_pcat.p, np, 1, flags.p);
            ++launches;
            launches += launch_exclusive_scan_u32_u64(s, flags.p, pos.p, np, tmp.p);
            CUDA_TRY(cudaMemcpyAsync(&K, pos.p + np, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
        }
        uint64_t Kf = 0, fkb = 0, R = 0;
        DevBuf<uint32_t> e_grp, e_sent, e_tok, perm_a, perm_b, keys_a, keys_b, hist;
        DevBuf<uint64_t> hist_off, stmp, flex_off, flex_ref_off;
        DevBuf<uint16_t> flex_len16;
        DevBuf<uint8_t>  flex_keys;
        DevBuf<uint32_t> flex_rs;
        DevBuf<uint16_t> flex_rt;
        uint32_t*        perm = nullptr;
        if (K) {
            TRY(sel.alloc(dev, K));
            TRY(lens.alloc(dev, K));
            TRY(sel_counts.alloc(dev, K));
            TRY(foff.alloc(dev, K + 1));
            scatter_selected_kernel<<<fx_div_up(np, 256), 256, 0, s>>>(flags.p, pos.p, np, sel.p);
            flex_lengths_kernel<<<fx_div_up(K, 256), 256, 0, s>>>(m->d_keys.p, m->d_off.p, sel.p, K, m->d_counts.p, lens.p, sel_counts.p);
            launches += 2;
            launches += launch_exclusive_scan_u32_u64(s, lens.p, foff.p, K, tmp.p);
            CUDA_TRY(cudaMemcpyAsync(&fkb, foff.p + K, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            TRY(fkeys.alloc(dev, fkb + 1));
            flex_write_kernel<<<fx_div_up(K, 256), 256, 0, s>>>(m->d_keys.p, m->d_off.p, sel.p, K, foff.p, fkeys.p);
            ++launches;
            // ---- distinct flexgrams: an index over the blob, duplicates resolved to the first claimant
            uint64_t cap = 1024, pbits = 1ull << 15;
            while (cap < 2 * K) cap <<= 1;
            while (pbits < 16 * K) pbits <<= 1;
            DevBuf<PatSlot>          slots;
            DevBuf<uint32_t>         presence;
            DevBuf<PatternMetaStats> st;
            PatternMetaStats         hst;
            TRY(slots.alloc(dev, cap));
            TRY(presence.alloc(dev, pbits / 32));
            TRY(st.alloc(dev, 1));
            TRY(rep.alloc(dev, K));
            CUDA_TRY(cudaMemsetAsync(slots.p, 0, cap * sizeof(PatSlot), s));
            CUDA_TRY(cudaMemsetAsync(presence.p, 0, pbits / 8, s));
            CUDA_TRY(cudaMemsetAsync(st.p, 0, sizeof(PatternMetaStats), s));
            launches += launch_index_build(s, fkeys.p, foff.p, K, slots.p, cap, presence.p, pbits, st.p, rep.p);
            CUDA_TRY(cudaMemcpyAsync(&hst, st.p, sizeof hst, cudaMemcpyDeviceToHost, s));
            TRY(uflags.alloc(dev, K + 1));
            TRY(upos.alloc(dev, K + 1));
            flag_representatives_kernel<<<fx_div_up(K, 256), 256, 0, s>>>(rep.p, K, uflags.p);
            ++launches;
            launches += launch_exclusive_scan_u32_u64(s, uflags.p, upos.p, K, tmp.p);
            CUDA_TRY(cudaMemcpyAsync(&Kf, upos.p + K, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            if (hst.malformed) return set_err(COLIBRI_E_CAPACITY, "flexgram index overflow");
            // ---- counts and occurrence triples
            TRY(uniq.alloc(dev, Kf));
            TRY(group.alloc(dev, K));
            TRY(flex_counts.alloc(dev, Kf));
            TRY(emit_off.alloc(dev, K + 1));
            CUDA_TRY(cudaMemsetAsync(flex_counts.p, 0, Kf * sizeof(uint32_t), s));
            scatter_selected_kernel<<<fx_div_up(K, 256), 256, 0, s>>>(uflags.p, upos.p, K, uniq.p);
            flex_groups_kernel<<<fx_div_up(K, 256), 256, 0, s>>>(rep.p, upos.p, sel_counts.p, K, group.p, flex_counts.p);
            launches += 2;
            launches += launch_exclusive_scan_u32_u64(s, sel_counts.p, emit_off.p, K, tmp.p);
            CUDA_TRY(cudaMemcpyAsync(&R, emit_off.p + K, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            if (R >= 0xFFFFFFF0ull) return set_err(COLIBRI_E_CAPACITY, "%llu skipgram occurrences; the device sort is 32 bit", (unsigned long long)R);
            // ---- the distinct flexgrams as a flat set
            TRY(flex_lens.alloc(dev, Kf));
            TRY(flex_len16.alloc(dev, Kf));
            TRY(flex_off.alloc(dev, Kf + 1));
            DevBuf<uint32_t> dummy_counts;
            TRY(dummy_counts.alloc(dev, Kf));
            launches += launch_gather_meta(s, uniq.p, Kf, foff.p, nullptr, nullptr, flex_lens.p, flex_len16.p, dummy_counts.p);
            launches += launch_exclusive_scan_u32_u64(s, flex_lens.p, flex_off.p, Kf, tmp.p);
            uint64_t ukb = 0;
            CUDA_TRY(cudaMemcpyAsync(&ukb, flex_off.p + Kf, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            TRY(flex_keys.alloc(dev, ukb + 1));
            launches += launch_gather_keys(s, uniq.p, Kf, fkeys.p, foff.p, flex_off.p, flex_keys.p);
            fkb = ukb;
            TRY(flex_ref_off.alloc(dev, Kf + 1));
            launches += launch_exclusive_scan_u32_u64(s, flex_counts.p, flex_ref_off.p, Kf, tmp.p);
            // ---- occurrences: emit (group, sentence, token), sort by token, sentence, group (stable LSD) -> ascending inside each flexgram
            TRY(flex_rs.alloc(dev, std::max<uint64_t>(R, 1)));
            TRY(flex_rt.alloc(dev, std::max<uint64_t>(R, 1)));
            if (R) {
                const uint64_t nsort = (R + 4095) / 4096;
                TRY(e_grp.alloc(dev, R));
                TRY(e_sent.alloc(dev, R));
                TRY(e_tok.alloc(dev, R));
                TRY(perm_a.alloc(dev, R));
                TRY(perm_b.alloc(dev, R));
                TRY(keys_a.alloc(dev, R));
                TRY(keys_b.alloc(dev, R));
                TRY(hist.alloc(dev, 256 * nsort));
                TRY(hist_off.alloc(dev, 256 * nsort + 1));
                TRY(stmp.alloc(dev, 256 * nsort / 2048 + 4));
                DevBuf<unsigned int> maxes;
                TRY(maxes.alloc(dev, 2));
                CUDA_TRY(cudaMemsetAsync(maxes.p, 0, 2 * sizeof(unsigned int), s));
                flex_emit_kernel<<<fx_div_up(K * 32, 256), 256, 0, s>>>(sel.p, K, group.p, emit_off.p, m->d_ref_off.p, m->d_ref_sentence.p, m->d_ref_token.p, e_grp.p, e_sent.p, e_tok.p);
                max_u32_kernel<<<1184, 256, 0, s>>>(e_sent.p, R, maxes.p);
                max_u32_kernel<<<1184, 256, 0, s>>>(e_tok.p, R, maxes.p + 1);
                launches += 3;
                unsigned int hmax[2] = {0, 0};
                CUDA_TRY(cudaMemcpyAsync(hmax, maxes.p, sizeof hmax, cudaMemcpyDeviceToHost, s));
                CUDA_TRY(cudaStreamSynchronize(s));
                launches += launch_iota(s, perm_a.p, R);
                perm = perm_a.p;
                uint32_t* alt = perm_b.p;
                TRY(sort_perm_by(s, e_tok.p, hmax[1], R, perm, alt, keys_a.p, keys_b.p, hist.p, hist_off.p, stmp.p, launches));
                TRY(sort_perm_by(s, e_sent.p, hmax[0], R, perm, alt, keys_a.p, keys_b.p, hist.p, hist_off.p, stmp.p, launches));
                TRY(sort_perm_by(s, e_grp.p, (uint32_t)(Kf ? Kf - 1 : 0), R, perm, alt, keys_a.p, keys_b.p, hist.p, hist_off.p, stmp.p, launches));
                gather_refs_perm_kernel<<<fx_div_up(R, 256), 256, 0, s>>>(e_sent.p, e_tok.p, perm, R, flex_rs.p, flex_rt.p);
                ++launches;
            }
        }
        // ---- the result: the old model with the flexgrams appended
        const uint64_t total = np + Kf;
        r->npatterns = total;
        r->keybytes  = m->keybytes + (Kf ? fkb : 0);
        r->nrefs     = m->nrefs + R;
        TRY(r->d_keys.alloc(dev, std::max<uint64_t>(r->keybytes, 1)));
        TRY(r->d_off.alloc(dev, total + 1));
        TRY(r->d_counts.alloc(dev, std::max<uint64_t>(total, 1)));
        TRY(r->d_len16.alloc(dev, std::max<uint64_t>(total, 1)));
        TRY(r->d_ref_off.alloc(dev, total + 1));
        TRY(r->d_ref_sentence.alloc(dev, std::max<uint64_t>(r->nrefs, 1)));
        TRY(r->d_ref_token.alloc(dev, std::max<uint64_t>(r->nrefs, 1)));
        if (m->keybytes) CUDA_TRY(cudaMemcpyAsync(r->d_keys.p, m->d_keys.p, m->keybytes, cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(r->d_off.p, m->d_off.p, (np + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(r->d_ref_off.p, m->d_ref_off.p, (np + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        if (np) {
            CUDA_TRY(cudaMemcpyAsync(r->d_counts.p, m->d_counts.p, np * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(r->d_len16.p, m->d_len16.p, np * sizeof(uint16_t), cudaMemcpyDeviceToDevice, s));
        }
        if (m->nrefs) {
            CUDA_TRY(cudaMemcpyAsync(r->d_ref_sentence.p, m->d_ref_sentence.p, m->nrefs * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(r->d_ref_token.p, m->d_ref_token.p, m->nrefs * sizeof(uint16_t), cudaMemcpyDeviceToDevice, s));
        }
        if (Kf) {
            CUDA_TRY(cudaMemcpyAsync(r->d_keys.p + m->keybytes, flex_keys.p, fkb, cudaMemcpyDeviceToDevice, s));
            shift_u64_kernel<<<fx_div_up(Kf + 1, 256), 256, 0, s>>>(flex_off.p, Kf + 1, m->keybytes, r->d_off.p + np);
            shift_u64_kernel<<<fx_div_up(Kf + 1, 256), 256, 0, s>>>(flex_ref_off.p, Kf + 1, m->nrefs, r->d_ref_off.p + np);
            launches += 2;
            CUDA_TRY(cudaMemcpyAsync(r->d_counts.p + np, flex_counts.p, Kf * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(r->d_len16.p + np, flex_len16.p, Kf * sizeof(uint16_t), cudaMemcpyDeviceToDevice, s));
            if (R) {
                CUDA_TRY(cudaMemcpyAsync(r->d_ref_sentence.p + m->nrefs, flex_rs.p, R * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
                CUDA_TRY(cudaMemcpyAsync(r->d_ref_token.p + m->nrefs, flex_rt.p, R * sizeof(uint16_t), cudaMemcpyDeviceToDevice, s));
            }
        }
        CUDA_TRY(cudaStreamSynchronize(s));  // the temporaries go back to the pool when this returns
        CUDA_TRY(cudaGetLastError());
        r->totaltokens  = m->totaltokens;
        r->totaltypes   = m->totaltypes;
        r->maxn         = m->maxn;
        r->minn         = m->minn;
        r->hasskipgrams = m->hasskipgrams;
        r->hasflexgrams = Kf ? 1 : m->hasflexgrams;
        r->passes       = m->passes;
        r->counters[2]  = launches;
        if (found) *found = Kf;
        return 0;
    };
    int rc = body();
    if (rc) {
        cudaStreamSynchronize(r->stream);
        colibri_b200_model_free(r);
        return rc;
    }
    *out = r;
    return 0;
}
