// Single-process multi-GPU context (ml_ctx_create_multi, include/machline_gpu.h): ONE host thread of the caller -- the
// reference's `program main` is a single process (src/main.f90:133) -- drives the two hot paths on several GPUs through the
// same entry points as a single-device context.  The handle is a facade over one ordinary context per device; every
// entry point fans out to them (one short-lived worker thread per device for the calls that launch kernels, because
// the sharded solvers are collective: all ranks must be inside ml_solve at the same time) and joins before it returns.
// The devices talk to each other exactly as the one-process-per-GPU ranks do: NCCL (ncclCommInitRank from each worker)
// and the peer-memory windows of the Krylov solvers, mapped here with cudaDeviceEnablePeerAccess instead of CUDA IPC
// (solve_kernels.cu: p2p_setup).  Rows of the permuted system are dealt to the devices in contiguous blocks, or
// block-cyclically after ml_multi_set_dealing (load balance of the sharded LU).
#include <algorithm>
#include <array>
#include <cstring>
#include <string>
#include <thread>

#include "ctx.h"

namespace mlgpu {

struct Group {
    std::vector<ml_ctx*> m;    // member contexts; rank = index
    int block_rows = -1;       // 0: contiguous row blocks; > 0: block-cyclic dealing with blocks of this many rows; -1 (default):
                               // contiguous for subsonic flows, block-cyclic 128 for supersonic ones, where the rows of a sorted
                               // system get more expensive downstream (measured on 8 GPUs, AGARD-B class: 26 vs 42 ms, 371 vs 624 ms)
    int n_cp = 0;
    int n_unknown = 0;
};

namespace {

// f(rank) on every member, concurrently; the first failing status is returned and its message copied to the facade.
template <class F>
ml_status for_all(ml_ctx* c, F f) {
    Group* g = c->group;
    const int n = (int)g->m.size();
    std::vector<ml_status> st(n, ML_OK);
    std::vector<std::thread> th;
    th.reserve(n > 0 ? n - 1 : 0);
    for (int i = 1; i < n; ++i) th.emplace_back([&, i] { st[i] = f(i); });
    st[0] = f(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (st[i] != ML_OK) {
            c->err = "device " + std::to_string(g->m[i]->device) + ": " + g->m[i]->err;
            return st[i];
        }
    return ML_OK;
}

template <class F>
ml_status for_each_seq(ml_ctx* c, F f) {
    Group* g = c->group;
    for (size_t i = 0; i < g->m.size(); ++i) {
        ml_status st = f((int)i);
        if (st != ML_OK) {
            c->err = "device " + std::to_string(g->m[i]->device) + ": " + g->m[i]->err;
            return st;
        }
    }
    return ML_OK;
}

ml_status deal_rows(ml_ctx* c) {
    Group* g = c->group;
    const int n = (int)g->m.size(), n_cp = g->n_cp;
    if (n_cp <= 0) return ML_OK;
    int block = g->block_rows;
    if (block < 0) block = (g->m[0]->have_flow && g->m[0]->flow.supersonic) ? 128 : 0;
    return for_each_seq(c, [&](int i) {
        if (block > 0) return ml_set_row_shard_cyclic(g->m[i], block, i, n);
        const int base = n_cp / n, rem = n_cp % n;
        return ml_set_row_shard(g->m[i], i * base + std::min(i, rem), base + (i < rem ? 1 : 0));
    });
}

// runs of consecutive global rows among a member's (ascending) local rows that fall inside [row0, row0 + nrows)
template <class F>
ml_status for_runs(ml_ctx* member, int row0, int nrows, F f) {
    const std::vector<int>& lr = member->local_rows;
    size_t i = std::lower_bound(lr.begin(), lr.end(), row0) - lr.begin();
    while (i < lr.size() && lr[i] < row0 + nrows) {
        size_t j = i + 1;
        while (j < lr.size() && lr[j] == lr[j - 1] + 1 && lr[j] < row0 + nrows) ++j;
        ml_status st = f(lr[i], (int)(j - i));
        if (st != ML_OK) return st;
        i = j;
    }
    return ML_OK;
}

}  // namespace

void multi_destroy(ml_ctx* c) {
    if (!c->group) return;
    for (ml_ctx* m : c->group->m) ml_ctx_destroy(m);
    delete c->group;
    c->group = nullptr;
}

// every member holds the replicated solution after a solve: the first one post-processes it
ml_status multi_post_process(ml_ctx* c, const ml_post_tables* t, const ml_post_flow* f, const double* x_override, ml_post_out* out) {
    ml_ctx* m0 = c->group->m[0];
    ml_status st = ml_post_process(m0, t, f, x_override, out);
    if (st != ML_OK) c->err = m0->err;
    return st;
}

ml_status multi_set_flow(ml_ctx* c, const ml_flow* f) {
    ml_status st = for_each_seq(c, [&](int i) { return ml_set_flow(c->group->m[i], f); });
    if (st != ML_OK) return st;
    c->assembled = false;
    return deal_rows(c);   // the automatic dealing depends on the flow regime
}
ml_status multi_set_panels(ml_ctx* c, const ml_panel_soa* body, const ml_panel_soa* wake) {
    return for_each_seq(c, [&](int i) { return ml_set_panels(c->group->m[i], body, wake); });
}
ml_status multi_set_system_map(ml_ctx* c, const ml_system_map* m) {
    if (m) c->group->n_unknown = m->n_unknown;
    return for_each_seq(c, [&](int i) { return ml_set_system_map(c->group->m[i], m); });
}
ml_status multi_set_control_points(ml_ctx* c, int n_cp, const double* loc, const int* bc, const double* n_g, const int* row_perm) {
    ml_status st = for_each_seq(c, [&](int i) { return ml_set_control_points(c->group->m[i], n_cp, loc, bc, n_g, row_perm); });
    if (st != ML_OK) return st;
    c->group->n_cp = n_cp;
    c->assembled = false;
    return deal_rows(c);
}
ml_status multi_set_dealing(ml_ctx* c, int block_rows) {
    c->group->block_rows = block_rows;
    c->assembled = false;
    return deal_rows(c);
}

ml_status multi_assemble(ml_ctx* c, double* I_known_out, bool resident, double* device_ms) {
    Group* g = c->group;
    std::vector<std::vector<double>> Ik(g->m.size());
    std::vector<double> ms(g->m.size(), 0.);
    ml_status st = for_all(c, [&](int i) {
        if (resident) return ml_assemble_resident(g->m[i], &ms[i]);
        ml_ctx* m = g->m[i];
        ml_status s = ml_assemble(m, nullptr);
        if (s == ML_OK) {
            Ik[i] = m->h_I_known;
            ms[i] = m->assemble_ms;
        }
        return s;
    });
    if (st != ML_OK) return st;
    c->assemble_ms = *std::max_element(ms.begin(), ms.end());
    if (device_ms) *device_ms = c->assemble_ms;
    c->assembled = true;
    if (!resident && I_known_out)
        for (size_t i = 0; i < g->m.size(); ++i)
            for (size_t k = 0; k < g->m[i]->local_rows.size(); ++k) I_known_out[g->m[i]->local_rows[k]] = Ik[i][k];
    return ML_OK;
}

ml_status multi_get_A(ml_ctx* c, int row0, int nrows, double* dst, int ld) {
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_get_A before ml_assemble");
    if (row0 < 0 || row0 + nrows > c->group->n_cp) return c->fail(ML_BAD_ARGUMENT, "rows out of range");
    return for_each_seq(c, [&](int i) {
        ml_ctx* m = c->group->m[i];
        return for_runs(m, row0, nrows, [&](int r0, int n) { return ml_get_A(m, r0, n, dst + (r0 - row0), ld); });
    });
}
ml_status multi_set_A(ml_ctx* c, int row0, int nrows, const double* src, int ld) {
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_set_A before ml_assemble");
    if (row0 < 0 || row0 + nrows > c->group->n_cp) return c->fail(ML_BAD_ARGUMENT, "rows out of range");
    return for_each_seq(c, [&](int i) {
        ml_ctx* m = c->group->m[i];
        return for_runs(m, row0, nrows, [&](int r0, int n) { return ml_set_A(m, r0, n, src + (r0 - row0), ld); });
    });
}

ml_status multi_local_rows(ml_ctx* c, int* rows_out, int* n_out) {
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_local_rows before ml_assemble");
    *n_out = c->group->n_cp;
    if (rows_out)
        for (int i = 0; i < c->group->n_cp; ++i) rows_out[i] = i;
    return ML_OK;
}

ml_status multi_solve(ml_ctx* c, const ml_solver_opts* opts, const double* BC, double* x_out, ml_solve_info* info) {
    Group* g = c->group;
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_solve before ml_assemble");
    const int n = (int)g->m.size();
    std::vector<std::vector<double>> x(n, std::vector<double>(g->n_unknown, 0.));
    std::vector<ml_solve_info> inf(n);
    ml_status st = for_all(c, [&](int i) { return ml_solve(g->m[i], opts, BC, x[i].data(), &inf[i]); });
    if (st != ML_OK) return st;
    std::memcpy(x_out, x[0].data(), sizeof(double) * g->n_unknown);   // every rank holds the full solution
    if (info) {
        *info = inf[0];
        for (int i = 1; i < n; ++i) {
            info->solve_ms = std::max(info->solve_ms, inf[i].solve_ms);
            info->assemble_ms = std::max(info->assemble_ms, inf[i].assemble_ms);
        }
    }
    return ML_OK;
}

ml_status multi_check_system(ml_ctx* c, const double* BC, int* n_zero_rows, int* n_zero_cols) {
    Group* g = c->group;
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_check_system before ml_assemble");
    const int n = (int)g->m.size();
    std::vector<int> zr(n, 0), zc(n, 0);
    std::vector<ml_status> verdict(n, ML_OK);
    // the verdict (0 / 1 / 2) is data, not a failure: collect it per rank (all ranks agree: the check all-reduces)
    ml_status st = for_all(c, [&](int i) {
        ml_status s = ml_check_system(g->m[i], BC, &zr[i], &zc[i]);
        if (s == ML_NAN_IN_SYSTEM || s == ML_UNINFLUENCED) {
            verdict[i] = s;
            return ML_OK;
        }
        return s;
    });
    if (st != ML_OK) return st;
    if (n_zero_rows) *n_zero_rows = zr[0];
    if (n_zero_cols) *n_zero_cols = zc[0];
    if (verdict[0] != ML_OK) return c->fail(verdict[0], g->m[0]->err);
    return ML_OK;
}

ml_status multi_residual(ml_ctx* c, const double* BC, const double* x, double* r_out) {
    Group* g = c->group;
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_residual before ml_assemble");
    std::vector<std::vector<double>> r(g->m.size());
    ml_status st = for_all(c, [&](int i) {
        r[i].assign(g->m[i]->local_rows.size() + 1, 0.);
        return ml_residual(g->m[i], BC, x, r[i].data());
    });
    if (st != ML_OK) return st;
    for (size_t i = 0; i < g->m.size(); ++i)
        for (size_t k = 0; k < g->m[i]->local_rows.size(); ++k) r_out[g->m[i]->local_rows[k]] = r[i][k];
    return ML_OK;
}

ml_status multi_dod_census(ml_ctx* c, long long* counts4) {
    Group* g = c->group;
    std::vector<std::array<long long, 4>> cnt(g->m.size());
    ml_status st = for_all(c, [&](int i) { return ml_dod_census(g->m[i], cnt[i].data()); });
    if (st != ML_OK) return st;
    for (int k = 0; k < 4; ++k) {
        counts4[k] = 0;
        for (auto& a : cnt) counts4[k] += a[k];
    }
    return ML_OK;
}

long long multi_sum_launches(const ml_ctx* c) {
    long long s = 0;
    for (ml_ctx* m : c->group->m) s += m->launches;
    return s;
}
long long multi_sum_pairs(const ml_ctx* c) {
    long long s = 0;
    for (ml_ctx* m : c->group->m) s += m->pair_count;
    return s;
}
ml_ctx* multi_member(const ml_ctx* c, int i) { return c->group->m[i]; }
int multi_size(const ml_ctx* c) { return (int)c->group->m.size(); }

ml_status multi_profile(ml_ctx* c, int what, int on, ml_profile* out) {   // what: 0 set, 1 get, 2 reset
    Group* g = c->group;
    if (what == 1) std::memset(out, 0, sizeof *out);
    return for_each_seq(c, [&](int i) {
        if (what == 0) return ml_set_profiling(g->m[i], on);
        if (what == 2) return ml_reset_profile(g->m[i]);
        ml_profile p;
        ml_status s = ml_get_profile(g->m[i], &p);
        if (s != ML_OK) return s;
        out->h2d_bytes += p.h2d_bytes;
        out->d2h_bytes += p.d2h_bytes;
        out->gemv_launches += p.gemv_launches;
        out->gemv_bytes += p.gemv_bytes;
        out->gemv_ms = std::max(out->gemv_ms, p.gemv_ms);
        out->assemble_ms = std::max(out->assemble_ms, p.assemble_ms);
        out->comm_ms = std::max(out->comm_ms, p.comm_ms);
        return ML_OK;
    });
}

}  // namespace mlgpu

extern "C" ml_status ml_ctx_create_multi(ml_ctx** out, const int* device_ids, int n_dev) {
    if (!out || !device_ids || n_dev < 1 || n_dev > mlgpu::Ctx::P2P_MAX) return ML_BAD_ARGUMENT;
    *out = nullptr;
    for (int i = 0; i < n_dev; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return ML_BAD_ARGUMENT;
    ml_ctx* c = new ml_ctx();
    c->group = new mlgpu::Group();
    c->device = device_ids[0];
    for (int i = 0; i < n_dev; ++i) {
        ml_ctx* m = nullptr;
        ml_status st = ml_ctx_create(&m, device_ids[i]);
        if (st != ML_OK) {
            ml_ctx_destroy(c);
            return st;
        }
        c->group->m.push_back(m);
    }
    if (n_dev > 1) {
        unsigned char id[128];
        ml_status st = ml_nccl_unique_id(id);
        if (st == ML_OK) st = mlgpu::for_all(c, [&](int i) { return ml_set_communicator(c->group->m[i], id, i, n_dev); });
        if (st != ML_OK) {
            ml_ctx_destroy(c);
            return st;
        }
    }
    *out = c;
    return ML_OK;
}

extern "C" ml_status ml_multi_set_dealing(ml_ctx* c, int block_rows) {
    if (!c || !c->group || block_rows < -1) return ML_BAD_ARGUMENT;
    return mlgpu::multi_set_dealing(c, block_rows);
}

extern "C" int ml_device_count(const ml_ctx* c) { return !c ? 0 : (c->group ? mlgpu::multi_size(c) : 1); }
