// cc_transition: the whole CrossCat kernel schedule behind one C call -- the shape of the reference's only
// native precedent, LocalEngine.analyze(M_c, T, X_L, X_D, seed, kernel_list, n_steps, ...)
// (src/crosscat/lovecat.py:377-391), and of State.transition / _transition_generic
// (src/crosscat/state.py:727-761,866-905) on the device random stream.  For every iteration and every kernel
// code, in the caller's order, over the listed chains, all rows and all columns:
//
//   CC_K_ROWS           View.transition_rows for every live view      (cc_transition_rows, with the capacity
//                       pause / resume protocol of rows_prog driven here)
//   CC_K_COLUMNS        State.transition_dims                          (cc_col_aux -> cc_col_marginals ->
//                       { cc_col_scan -> [cc_col_aux] -> cc_col_install -> cc_col_marginals(new views) }* ->
//                       cc_rebuild(mask): state.py:804-810,1050-1154)
//   CC_K_ALPHA          column-CRP concentration                       (state.py:763-765)
//   CC_K_VIEW_ALPHAS    row-CRP concentration of every view            (state.py:767-772)
//   CC_K_COLUMN_HYPERS  hyperparameters of every column                (state.py:781-786)
//
// When CC_K_COLUMNS directly follows CC_K_ROWS the auxiliary views of the column sweep -- which depend on the
// data, the hyperparameters and their own random keys only -- are generated on a side stream beside the
// latency-bound rows sweep.  The only host round trips are the small reads a data-dependent schedule needs:
// the views' column counts per kernel, and one read of the chains' pending births per round of the column scan.
// Capacity (cluster slots, view slots) is the caller's memory: when it runs out the `grow` callback re-allocates
// and updates the cc_state in place, and the schedule continues where it stopped.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "cc_common.cuh"

namespace {

struct Small {                       // host mirror of the small per-chain arrays
    std::vector<int32_t> nv, zv, prog;
};

int read_small(const cc_state* st, Small& m, bool with_prog, cudaStream_t stream) {
    m.nv.resize((size_t)st->S * st->Vcap);
    m.zv.resize((size_t)st->S * st->C);
    CC_CUDA(cudaMemcpyAsync(m.nv.data(), st->nv, m.nv.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CC_CUDA(cudaMemcpyAsync(m.zv.data(), st->zv, m.zv.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    if (with_prog) {
        m.prog.resize((size_t)st->S * st->Vcap * 4);
        CC_CUDA(cudaMemcpyAsync(m.prog.data(), st->rows_prog, m.prog.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    }
    CC_CUDA(cudaStreamSynchronize(stream));
    return CC_OK;
}

// splitmix64: column visiting orders (rng.permutation(outputs), state.py:807, on the device stream's seed)
inline uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// host copy of the device generator (cc_common.cuh: Philox), first output word: orders the auxiliary views by cost
inline uint32_t philox_first(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t seed) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ b, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return c0;
}

struct Ctx {
    cc_state* st;
    std::vector<int32_t> chains;
    const int32_t *ci_off, *ci_adj;
    uint64_t seed;
    uint64_t* counter;
    cc_grow_fn grow;
    void* user;
    cudaStream_t stream, side;
    cudaEvent_t ev_fork, ev_aux;
};

int need_room(Ctx& x, int min_kcap, int min_vcap, const char* what) {
    if (!x.grow) {
        cc_set_error("cc_transition: %s needs Kcap >= %d, Vcap >= %d (have %d, %d) and no grow callback was given",
                     what, min_kcap, min_vcap, x.st->Kcap, x.st->Vcap);
        return CC_ERR_CAPACITY;
    }
    CC_CUDA(cudaStreamSynchronize(x.stream));
    if (x.grow(x.user, min_kcap, min_vcap) != 0 || x.st->Kcap < min_kcap || x.st->Vcap < min_vcap) {
        cc_set_error("cc_transition: the grow callback did not provide Kcap >= %d, Vcap >= %d for %s", min_kcap, min_vcap, what);
        return CC_ERR_CAPACITY;
    }
    return CC_OK;
}

// ---- rows -----------------------------------------------------------------------------------------------
int do_rows(Ctx& x) {
    const cc_state* st = x.st;
    Small m;
    if (int rc = read_small(st, m, true, x.stream)) return rc;
    std::vector<int32_t> ctype(st->C);
    CC_CUDA(cudaMemcpyAsync(ctype.data(), st->ctype, sizeof(int32_t) * st->C, cudaMemcpyDeviceToHost, x.stream));
    CC_CUDA(cudaStreamSynchronize(x.stream));
    std::vector<cc_row_work> work;
    for (int s : x.chains)
        for (int v = 0; v < st->Vcap; ++v) {
            if (m.nv[(size_t)s * st->Vcap + v] <= 0) continue;
            cc_row_work w;
            w.chain = s; w.view = v; w.n = st->R; w.perm_is_index = 1; w.perm_off = -1; w.u_off = -1; w.trace_off = -1;
            w.n_cols = m.nv[(size_t)s * st->Vcap + v];
            int nn = 0;
            for (int c = 0; c < st->C; ++c) nn += (m.zv[(size_t)s * st->C + c] == v && ctype[c] == CC_NORMAL);
            w.n_normal = nn; w.resume = 0;
            w.cost_hint = m.prog[((size_t)s * st->Vcap + v) * 4 + 3];
            work.push_back(w);
        }
    if (work.empty() || st->R == 1) return CC_OK;
    const uint64_t ctr = ++*x.counter;
    int rc = cc_transition_rows(x.st, (int)work.size(), work.data(), nullptr, nullptr, nullptr, 0, x.seed, ctr, x.stream);
    if (rc) return rc;
    return CC_OK;
}

// complete a rows sweep: while some view stopped for lack of cluster slots, grow and resume (rows_prog)
int finish_rows(Ctx& x, uint64_t ctr) {
    while (x.st->Kcap < x.st->R + 1) {
        Small m;
        if (int rc = read_small(x.st, m, true, x.stream)) return rc;
        std::vector<cc_row_work> again;
        for (int s : x.chains)
            for (int v = 0; v < x.st->Vcap; ++v) {
                const int32_t* pr = &m.prog[((size_t)s * x.st->Vcap + v) * 4];
                if (m.nv[(size_t)s * x.st->Vcap + v] > 0 && pr[2] == CC_ROWS_GROW) {
                    cc_row_work w;
                    w.chain = s; w.view = v; w.n = x.st->R; w.perm_is_index = 1; w.perm_off = -1; w.u_off = -1; w.trace_off = -1;
                    w.n_cols = m.nv[(size_t)s * x.st->Vcap + v]; w.n_normal = 0; w.resume = 1; w.cost_hint = 0;
                    again.push_back(w);
                }
            }
        if (again.empty()) break;
        if (int rc = need_room(x, std::min(2 * x.st->Kcap, x.st->R + 1), x.st->Vcap, "a rows sweep")) return rc;
        if (int rc = cc_transition_rows(x.st, (int)again.size(), again.data(), nullptr, nullptr, nullptr, 0, x.seed, ctr, x.stream)) return rc;
    }
    return CC_OK;
}

// ---- columns ----------------------------------------------------------------------------------------------
struct ColPlan {
    uint64_t ctr;
    std::vector<int32_t> perm;          // [n_chains][C]
    bool stored, aux_launched;
};

void plan_columns(Ctx& x, ColPlan& p, uint64_t ctr) {
    const int C = x.st->C, n = (int)x.chains.size();
    p.ctr = ctr;
    p.perm.resize((size_t)n * C);
    for (int i = 0; i < n; ++i) {
        uint64_t sd = x.seed * 0x9E3779B97F4A7C15ull + ctr * 1009 + (uint64_t)x.chains[i];
        int32_t* q = &p.perm[(size_t)i * C];
        for (int c = 0; c < C; ++c) q[c] = c;
        for (int c = C - 1; c > 0; --c) { const int j = (int)(splitmix(sd) % (uint64_t)(c + 1)); std::swap(q[c], q[j]); }
    }
    p.stored = x.st->aux_zr_all != nullptr;
    p.aux_launched = false;
}

int launch_aux(Ctx& x, ColPlan& p, cudaStream_t strm) {
    const int C = x.st->C, n = (int)x.chains.size();
    std::vector<int32_t> a_ch((size_t)n * C), a_co((size_t)n * C), order((size_t)n * C);
    std::vector<uint32_t> key((size_t)n * C);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < C; ++j) {
            const size_t e = (size_t)i * C + j;
            a_ch[e] = x.chains[i]; a_co[e] = p.perm[e]; order[e] = (int32_t)e;
            // the pair's alpha index on the device stream (columns.cu: stream 4): the cost of a pair grows with it
            const uint64_t sid = ((uint64_t)4 << 56) | ((uint64_t)(a_ch[e] & 0xFFFFFF) << 32) | (uint32_t)a_co[e];
            key[e] = (uint32_t)(((uint64_t)philox_first(p.ctr, sid, x.seed) * CC_NGRID) >> 32);
        }
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return key[a] > key[b]; });
    std::vector<int32_t> s_ch(order.size()), s_co(order.size());
    for (size_t e = 0; e < order.size(); ++e) { s_ch[e] = a_ch[order[e]]; s_co[e] = a_co[order[e]]; }
    p.aux_launched = true;
    return cc_col_aux(x.st, (int)s_ch.size(), s_ch.data(), s_co.data(), nullptr, nullptr, p.stored ? 2 : 0, x.seed, p.ctr, strm);
}

int do_columns(Ctx& x, ColPlan& p) {
    cc_state* st = x.st;
    const int C = st->C, n = (int)x.chains.size();
    if (!st->arena) { cc_set_error("cc_transition: CC_K_COLUMNS needs cc_state.arena (see cc_col_aux)"); return CC_ERR_ARG; }
    if (p.aux_launched) CC_CUDA(cudaStreamWaitEvent(x.stream, x.ev_aux, 0));
    else if (int rc = launch_aux(x, p, x.stream)) return rc;
    Small m;
    if (int rc = read_small(st, m, false, x.stream)) return rc;
    std::vector<int32_t> w_ch, w_vw;
    for (int s : x.chains)
        for (int v = 0; v < st->Vcap; ++v)
            if (m.nv[(size_t)s * st->Vcap + v] > 0) { w_ch.push_back(s); w_vw.push_back(v); }
    if (int rc = cc_col_marginals(st, (int)w_ch.size(), w_ch.data(), w_vw.data(), nullptr, nullptr, nullptr, C, x.stream)) return rc;
    int32_t *ch_d = nullptr, *cols_d = nullptr, *pos_d = nullptr, *pend_d = nullptr;
    CC_CUDA(cudaMallocAsync(&ch_d, sizeof(int32_t) * n, x.stream));
    CC_CUDA(cudaMallocAsync(&cols_d, sizeof(int32_t) * (size_t)n * C, x.stream));
    CC_CUDA(cudaMallocAsync(&pos_d, sizeof(int32_t) * n, x.stream));
    CC_CUDA(cudaMallocAsync(&pend_d, sizeof(int32_t) * n, x.stream));
    CC_CUDA(cudaMemcpyAsync(ch_d, x.chains.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, x.stream));
    CC_CUDA(cudaMemcpyAsync(cols_d, p.perm.data(), sizeof(int32_t) * (size_t)n * C, cudaMemcpyHostToDevice, x.stream));
    CC_CUDA(cudaMemsetAsync(pos_d, 0, sizeof(int32_t) * n, x.stream));
    CC_CUDA(cudaMemsetAsync(pend_d, 0xff, sizeof(int32_t) * n, x.stream));
    std::vector<int32_t> pend(n), pos(n);
    int rc = CC_OK;
    while (rc == CC_OK) {
        rc = cc_col_scan(st, n, ch_d, cols_d, C, pos_d, pend_d, nullptr, x.seed, p.ctr, x.ci_off, x.ci_adj, nullptr, 0, x.stream);
        if (rc) break;
        CC_CUDA(cudaMemcpyAsync(pend.data(), pend_d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, x.stream));
        CC_CUDA(cudaMemcpyAsync(pos.data(), pos_d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, x.stream));
        CC_CUDA(cudaStreamSynchronize(x.stream));
        std::vector<int32_t> idx, b_ch, b_co;
        for (int i = 0; i < n; ++i) if (pend[i] >= 0) { idx.push_back(i); b_ch.push_back(x.chains[i]); b_co.push_back(p.perm[(size_t)i * C + pend[i]]); }
        if (idx.empty()) break;
        if (!p.stored) {       // regenerate the same auxiliary partitions (same keys) and keep them per chain
            rc = cc_col_aux(st, (int)b_ch.size(), b_ch.data(), b_co.data(), nullptr, nullptr, 1, x.seed, p.ctr, x.stream);
            if (rc) break;
        }
        // room for the births: cluster slots for the largest partition, a free view slot in every chain
        std::vector<int32_t> auxK(p.stored ? (size_t)st->S * C : (size_t)st->S);
        CC_CUDA(cudaMemcpyAsync(auxK.data(), p.stored ? st->aux_K_all : st->aux_K, auxK.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, x.stream));
        if ((rc = read_small(st, m, false, x.stream))) break;
        int kmax = 0;
        bool full = false;
        for (size_t b = 0; b < b_ch.size(); ++b) {
            kmax = std::max(kmax, p.stored ? auxK[(size_t)b_ch[b] * C + b_co[b]] : auxK[b_ch[b]]);
            bool any_free = false;
            for (int v = 0; v < st->Vcap; ++v) any_free |= m.nv[(size_t)b_ch[b] * st->Vcap + v] <= 0;
            full |= !any_free;
        }
        if (kmax > st->Kcap - 1 || full) {
            rc = need_room(x, kmax > st->Kcap - 1 ? kmax + 1 + std::max(16, kmax / 8) : st->Kcap,
                           full ? std::min(st->C, std::max(st->Vcap + 8, 3 * st->Vcap / 2)) : st->Vcap, "a view birth");
            if (rc || (rc = read_small(st, m, false, x.stream))) break;
        }
        if ((rc = cc_col_install(st, (int)b_ch.size(), b_ch.data(), b_co.data(), p.stored ? 2 : 1, x.stream))) break;
        std::vector<int32_t> k_ch, k_vw, k_off, k_cnt;
        int mx = 0;
        for (size_t b = 0; b < idx.size(); ++b) {
            const int i = idx[b], rem = C - pend[i] - 1;
            int slot = 0;
            for (int v = 0; v < st->Vcap; ++v) if (m.nv[(size_t)b_ch[b] * st->Vcap + v] <= 0) { slot = v; break; }
            if (rem > 0) { k_ch.push_back(b_ch[b]); k_vw.push_back(slot); k_off.push_back(i * C + pend[i] + 1); k_cnt.push_back(rem); mx = std::max(mx, rem); }
            pos[i] += 1;
        }
        if (!k_ch.empty() && (rc = cc_col_marginals(st, (int)k_ch.size(), k_ch.data(), k_vw.data(), cols_d, k_off.data(), k_cnt.data(), mx, x.stream))) break;
        CC_CUDA(cudaMemcpyAsync(pos_d, pos.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, x.stream));
        CC_CUDA(cudaStreamSynchronize(x.stream));
    }
    cudaFreeAsync(ch_d, x.stream); cudaFreeAsync(cols_d, x.stream); cudaFreeAsync(pos_d, x.stream); cudaFreeAsync(pend_d, x.stream);
    if (rc) return rc;
    // re-incorporate every visited column under its final view (state.py:1110-1121)
    std::vector<uint8_t> mask((size_t)st->S * C, 0);
    for (int s : x.chains) std::fill(mask.begin() + (size_t)s * C, mask.begin() + (size_t)(s + 1) * C, (uint8_t)1);
    uint8_t* mask_d = nullptr;
    CC_CUDA(cudaMallocAsync(&mask_d, mask.size(), x.stream));
    CC_CUDA(cudaMemcpyAsync(mask_d, mask.data(), mask.size(), cudaMemcpyHostToDevice, x.stream));
    if ((rc = read_small(st, m, false, x.stream))) return rc;
    w_ch.clear(); w_vw.clear();
    for (int s : x.chains)
        for (int v = 0; v < st->Vcap; ++v)
            if (m.nv[(size_t)s * st->Vcap + v] > 0) { w_ch.push_back(s); w_vw.push_back(v); }
    rc = cc_rebuild(st, (int)w_ch.size(), w_ch.data(), w_vw.data(), mask_d, x.stream);
    cudaFreeAsync(mask_d, x.stream);
    return rc;
}

}  // namespace

extern "C" int cc_transition(cc_state* st, int n_chains, const int32_t* chains, int n_iters, int n_kernels,
                             const int32_t* kernels, const int32_t* ci_off_dev, const int32_t* ci_adj_dev,
                             uint64_t seed, uint64_t* counter, cc_grow_fn grow, void* user, void* stream_) {
    if (!st || n_chains < 0 || n_iters < 0 || n_kernels < 0 || !counter || (n_chains > 0 && !chains) || (n_kernels > 0 && !kernels)) {
        cc_set_error("cc_transition: bad arguments");
        return CC_ERR_ARG;
    }
    for (int k = 0; k < n_kernels; ++k)
        if (kernels[k] < CC_K_ALPHA || kernels[k] > CC_K_COLUMNS) { cc_set_error("cc_transition: unknown kernel code %d", kernels[k]); return CC_ERR_ARG; }
    for (int i = 0; i < n_chains; ++i)
        if (chains[i] < 0 || chains[i] >= st->S) { cc_set_error("cc_transition: chain %d out of range", chains[i]); return CC_ERR_ARG; }
    if (n_chains == 0 || n_iters == 0 || n_kernels == 0) return CC_OK;
    cc_prepare_pool();
    int dev = 0;
    CC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { cc_set_error("cc_transition: device ordinal %d not supported", dev); return CC_ERR_ARG; }
    static cudaStream_t side[64] = {nullptr};
    static cudaEvent_t ev_fork[64] = {nullptr}, ev_aux[64] = {nullptr};
    if (!side[dev]) {
        CC_CUDA(cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking));
        CC_CUDA(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
        CC_CUDA(cudaEventCreateWithFlags(&ev_aux[dev], cudaEventDisableTiming));
    }
    Ctx x;
    x.st = st; x.chains.assign(chains, chains + n_chains); x.ci_off = ci_off_dev; x.ci_adj = ci_adj_dev;
    x.seed = seed; x.counter = counter; x.grow = grow; x.user = user;
    x.stream = (cudaStream_t)stream_; x.side = side[dev]; x.ev_fork = ev_fork[dev]; x.ev_aux = ev_aux[dev];
    const bool overlap = getenv("CGPM_B200_NO_AUX_OVERLAP") == nullptr;
    for (int it = 0; it < n_iters; ++it)
        for (int k = 0; k < n_kernels; ++k) {
            int rc = CC_OK;
            const cc_state* cst = st;
            switch (kernels[k]) {
            case CC_K_ROWS: {
                const bool then_cols = overlap && k + 1 < n_kernels && kernels[k + 1] == CC_K_COLUMNS && st->arena;
                if (then_cols) CC_CUDA(cudaEventRecord(x.ev_fork, x.stream));
                rc = do_rows(x);
                const uint64_t rows_ctr = *x.counter;
                ColPlan plan;
                if (rc == CC_OK && then_cols) {
                    // the next kernel's auxiliary views, beside the rows sweep
                    plan_columns(x, plan, ++*x.counter);
                    CC_CUDA(cudaStreamWaitEvent(x.side, x.ev_fork, 0));
                    rc = launch_aux(x, plan, x.side);
                    CC_CUDA(cudaEventRecord(x.ev_aux, x.side));
                }
                if (rc == CC_OK) rc = finish_rows(x, rows_ctr);
                if (rc == CC_OK && then_cols) { rc = do_columns(x, plan); ++k; }
                break;
            }
            case CC_K_COLUMNS: {
                ColPlan plan;
                plan_columns(x, plan, ++*x.counter);
                rc = do_columns(x, plan);
                break;
            }
            case CC_K_ALPHA:
                rc = cc_transition_alphas(cst, n_chains, chains, nullptr, nullptr, 1, seed, ++*x.counter, x.stream);
                break;
            case CC_K_VIEW_ALPHAS: {
                Small m;
                if ((rc = read_small(st, m, false, x.stream))) break;
                std::vector<int32_t> w_ch, w_vw;
                for (int s : x.chains)
                    for (int v = 0; v < st->Vcap; ++v)
                        if (m.nv[(size_t)s * st->Vcap + v] > 0) { w_ch.push_back(s); w_vw.push_back(v); }
                rc = cc_transition_alphas(cst, (int)w_ch.size(), w_ch.data(), w_vw.data(), nullptr, 2, seed, ++*x.counter, x.stream);
                break;
            }
            case CC_K_COLUMN_HYPERS: {
                std::vector<int32_t> h_ch, h_co;
                for (int s : x.chains) for (int c = 0; c < st->C; ++c) { h_ch.push_back(s); h_co.push_back(c); }
                rc = cc_transition_col_hypers(cst, (int)h_ch.size(), h_ch.data(), h_co.data(), nullptr, nullptr, seed, ++*x.counter, x.stream);
                break;
            }
            }
            if (rc) return rc;
        }
    return CC_OK;
}
