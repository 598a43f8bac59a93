// bprmf_step.cu -- one whole BPRMF training step enqueued from C (b2r_bprmf_train_step) on a step context.
//
// Stream plan (step t):
//   main :  fused gather+score+loss+dQ (rows read once) -> mean(loss) -> [wait plan(t)] -> segment+opt(I) -> segment+opt(U)
//   side :  plan(t+1) for the NEXT batch's ids (or plan(t) itself when nothing was prefetched)
// The index plan (radix sort of the ids) only depends on the ids, so the context builds it one step ahead on
// a library-owned non-blocking stream while the HBM-bound kernels of the current step run; joins are event
// waits, never host synchronisations.  Plans (bucket partitions, bucket.cu) are double-buffered in the caller-provided
// workspace.  d must be 32, 64 or 128 (the widths the bucket kernels are built for).
#include <stdlib.h>

#include "common.cuh"
#include "plan_direct.cuh"

extern "C" int b2r_bprmf_fused_fwd_bwd(const float* U, const int64_t* uid, int64_t n_users, const float* I,
                                       const int64_t* iid, int64_t n_items, float* pred, float* grad_pred,
                                       float* row_loss, float* dQ, int B, int C, int d, int32_t* err_flag,
                                       b2r_stream_t stream);

int b2r_bprmf_fused_fwd_bwd_loss(const float* U, const int64_t* uid, int64_t n_users, const float* I, const int64_t* iid,
                                 int64_t n_items, float* grad_pred, float* row_loss, float* dQ, float* qout, int B, int C,
                                 int d, int32_t* err_flag, float* loss_out, unsigned int* done_counter,
                                 b2r_stream_t stream);

int b2r_bprmf_flash_launch(const float* U, const int64_t* uid, int64_t n_users, const float* I, const int64_t* iid,
                           int64_t n_items, float* pred, float* grad_pred, float* row_loss, float* dQ, float* qout,
                           int B, int C, int d, int32_t* err_flag, float* loss_out, unsigned int* done_counter,
                           const b2r::DirectPlanDev* plan_i, const b2r::DirectPlanDev* plan_u, b2r_stream_t stream);

namespace b2r {

struct PlanBuf {
    size_t iws, uws;                 // bucket workspaces (b2r_bucket_partition) for the item / user table
};

struct StepLayout {
    size_t q, pred, g, rows, dQ, counter;
    PlanBuf plan[2];
    size_t iws_bytes, uws_bytes;
    size_t dir_i[2], dir_u[2], dir_i_bytes, dir_u_bytes; // direct plans (plan_direct.cuh), double-buffered; 0: not available
    size_t total;
};

static bool step_layout(int B, int C, int d, int64_t n_users, int64_t n_items, StepLayout* L) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align_up(bytes, 256);
        return o;
    };
    const size_t n = (size_t)B * C;
    L->q = take((size_t)B * d * 4);
    L->pred = take(n * 4);
    L->g = take(n * 4);
    L->rows = take((size_t)B * 4);
    L->dQ = take((size_t)B * d * 4);
    L->counter = take(256);
    L->iws_bytes = b2r_bucket_workspace_bytes((int64_t)n, n_items);
    L->uws_bytes = b2r_bucket_workspace_bytes(B, n_users);
    for (int s = 0; s < 2; ++s) {
        L->plan[s].iws = take(L->iws_bytes);
        L->plan[s].uws = take(L->uws_bytes);
    }
    L->dir_i_bytes = direct_workspace_bytes((int64_t)n, n_items);
    L->dir_u_bytes = direct_workspace_bytes(B, n_users);
    if (L->dir_i_bytes == 0 || L->dir_u_bytes == 0) L->dir_i_bytes = L->dir_u_bytes = 0;
    for (int sl = 0; sl < 2; ++sl) {
        L->dir_i[sl] = take(L->dir_i_bytes);
        L->dir_u[sl] = take(L->dir_u_bytes);
    }
    L->total = off;
    return L->iws_bytes != 0 && L->uws_bytes != 0;
}

struct StepCtx {
    int B, C, d;
    int64_t n_users, n_items;
    StepLayout L;
    char* ws;
    cudaStream_t side;                // builds the plan of the next batch
    cudaEvent_t fork, join[2];
    int slot;                         // plan buffer the NEXT step will read if it was prefetched
    const void* pre_uid;
    const void* pre_iid;
    bool have_pre;
    int mode;                         // 0: direct plans built one batch ahead on the side stream (default);
                                      // 1: the forward kernel fills the direct plans itself, one stream (B2R_STEP=inline);
                                      // 2: count / scan / scatter / sort bucket plans one batch ahead (B2R_STEP=legacy)
};

}  // namespace b2r

using namespace b2r;

extern "C" size_t b2r_bprmf_step_workspace_bytes(int B, int C, int d, int64_t n_users, int64_t n_items) {
    if (B <= 0 || C <= 0 || d <= 0) return 0;
    StepLayout L;
    if (!step_layout(B, C, d, n_users, n_items, &L)) return 0;
    return L.total;
}

extern "C" int b2r_bprmf_ctx_create(void** ctx_out, int B, int C, int d, int64_t n_users, int64_t n_items, void* ws,
                                    size_t ws_bytes) {
    B2R_REQUIRE(ctx_out && ws, B2R_E_BADARG, "b2r_bprmf_ctx_create: null pointer");
    B2R_REQUIRE(B > 0 && C > 0 && d > 0 && d % 4 == 0, B2R_E_BADARG, "b2r_bprmf_ctx_create: B=%d C=%d d=%d", B, C, d);
    B2R_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, B2R_E_BADARG, "workspace must be 256-byte aligned");
    StepCtx* c = new StepCtx();
    c->B = B; c->C = C; c->d = d; c->n_users = n_users; c->n_items = n_items;
    if (!step_layout(B, C, d, n_users, n_items, &c->L)) {
        delete c;
        return set_error(B2R_E_UNSUPPORTED, "b2r_bprmf_ctx_create: cannot size the index plans");
    }
    if (ws_bytes < c->L.total) {
        const size_t need = c->L.total;
        delete c;
        return set_error(B2R_E_WORKSPACE, "b2r_bprmf_ctx_create: workspace %zu < required %zu", ws_bytes, need);
    }
    c->ws = static_cast<char*>(ws);
    c->slot = 0;
    c->have_pre = false;
    c->pre_uid = c->pre_iid = nullptr;
    // The plan stream gets the highest priority: its kernels are small, and when the HBM-bound kernels of the main
    // stream occupy every SM the plan's CTAs must be placed first whenever a slot frees up -- otherwise the plan of
    // the next batch is starved and the next step's update waits for it (measured: 0.177 vs 0.199 ms per step).
    int prio_lo = 0, prio_hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    // B2R_SIDE_PRIO=low (A/B knob): plan stream at the lowest priority instead
    const char* sp = getenv("B2R_SIDE_PRIO");
    const int side_prio = (sp && sp[0] == 'l') ? prio_lo : prio_hi;
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, side_prio);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->join[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->join[1], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete c;
        return set_error((int)e, "b2r_bprmf_ctx_create: %s", cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(c->ws + c->L.counter, 0, 256, c->side);
    if (e != cudaSuccess) {
        delete c;
        return set_error((int)e, "b2r_bprmf_ctx_create: %s", cudaGetErrorString(e));
    }
    for (int sl = 0; sl < 2; ++sl) {      // bucket counters start at zero (the scan kernel re-zeroes them every step)
        int rc = b2r_bucket_workspace_init(c->ws + c->L.plan[sl].iws, c->L.iws_bytes, (int64_t)B * C, n_items, c->side);
        if (rc == 0) rc = b2r_bucket_workspace_init(c->ws + c->L.plan[sl].uws, c->L.uws_bytes, B, n_users, c->side);
        if (rc != 0) {
            delete c;
            return rc;
        }
    }
    // Index plans.  Default (mode 0): "direct" plans (plan_direct.cuh: one scatter launch into fixed-capacity bucket regions,
    // one sort launch, both tables per launch) built for the NEXT batch on the high-priority side stream underneath this
    // step's kernels.  B2R_STEP=inline (mode 1): the forward kernel scatters the pairs itself, everything on one stream.
    // B2R_STEP=legacy (mode 2): the round-1 form, count / scan / scatter / sort per table (8 launches) one batch ahead.
    const char* mode = getenv("B2R_STEP");
    c->mode = (mode && mode[0] == 'i') ? 1 : ((mode && mode[0] == 'l') ? 2 : 0);
    if (c->L.dir_i_bytes == 0) c->mode = 2;
    if (c->mode != 2) {
        for (int sl = 0; sl < 2; ++sl) {
            int rc = direct_workspace_init(c->ws + c->L.dir_i[sl], c->L.dir_i_bytes, (int64_t)B * C, n_items, c->side);
            if (rc == 0) rc = direct_workspace_init(c->ws + c->L.dir_u[sl], c->L.dir_u_bytes, B, n_users, c->side);
            if (rc != 0) {
                delete c;
                return rc;
            }
        }
    }
    e = cudaStreamSynchronize(c->side);
    if (e != cudaSuccess) {
        delete c;
        return set_error((int)e, "b2r_bprmf_ctx_create: %s", cudaGetErrorString(e));
    }
    *ctx_out = c;
    return 0;
}

extern "C" int b2r_bprmf_ctx_destroy(void* ctx) {
    if (!ctx) return 0;
    StepCtx* c = static_cast<StepCtx*>(ctx);
    cudaStreamSynchronize(c->side);
    cudaStreamDestroy(c->side);
    cudaEventDestroy(c->fork);
    cudaEventDestroy(c->join[0]);
    cudaEventDestroy(c->join[1]);
    delete c;
    return 0;
}

extern "C" int b2r_bprmf_ctx_reset(void* ctx) {
    if (!ctx) return 0;
    StepCtx* c = static_cast<StepCtx*>(ctx);
    c->have_pre = false;
    c->pre_uid = c->pre_iid = nullptr;
    return 0;
}

static int build_plans(StepCtx* c, int slot, const int64_t* uid, const int64_t* iid, int32_t* err_flag) {
    char* base = c->ws;
    if (c->mode == 0) {
        const int64_t nn = (int64_t)c->B * c->C;
        profile_begin(B2R_PROF_PLAN_I, c->side);
        int rc = direct_scatter_pair(iid, nn, c->n_items, base + c->L.dir_i[slot], uid, c->B, c->n_users,
                                     base + c->L.dir_u[slot], err_flag, c->side);
        if (rc == 0) rc = direct_sort_pair(base + c->L.dir_i[slot], nn, c->n_items, base + c->L.dir_u[slot], c->B, c->n_users, c->side);
        if (rc != 0) return rc;
        profile_end(B2R_PROF_PLAN_I, c->side);
        B2R_CUDA_OK(cudaEventRecord(c->join[slot], c->side));
        return 0;
    }
    const PlanBuf& p = c->L.plan[slot];
    const int64_t n = (int64_t)c->B * c->C;
    profile_begin(B2R_PROF_PLAN_I, c->side);
    int rc = b2r_bucket_partition(iid, n, c->n_items, -1, 0, base + p.iws, c->L.iws_bytes, err_flag, c->side);
    if (rc != 0) return rc;
    profile_end(B2R_PROF_PLAN_I, c->side);
    rc = b2r_bucket_partition(uid, c->B, c->n_users, -1, 0, base + p.uws, c->L.uws_bytes, err_flag, c->side);
    if (rc != 0) return rc;
    B2R_CUDA_OK(cudaEventRecord(c->join[slot], c->side));
    return 0;
}

extern "C" int b2r_bprmf_train_step(void* ctx, const b2r_bprmf_tables* t, const int64_t* uid, const int64_t* iid,
                                    const int64_t* next_uid, const int64_t* next_iid, const b2r_optim* opt,
                                    float* loss_out, int32_t* err_flag, b2r_stream_t stream) {
    B2R_REQUIRE(ctx && t && uid && iid && opt && loss_out, B2R_E_BADARG, "b2r_bprmf_train_step: null pointer");
    StepCtx* c = static_cast<StepCtx*>(ctx);
    B2R_REQUIRE(t->d == c->d && t->n_users == c->n_users && t->n_items == c->n_items, B2R_E_BADARG,
                "b2r_bprmf_train_step: tables do not match the context");
    const int B = c->B, C = c->C, d = c->d;
    char* base = c->ws;
    float* pred = reinterpret_cast<float*>(base + c->L.pred);
    float* g = reinterpret_cast<float*>(base + c->L.g);
    float* rows = reinterpret_cast<float*>(base + c->L.rows);
    float* dQ = reinterpret_cast<float*>(base + c->L.dQ);
    cudaStream_t main_s = as_stream(stream);
    const int64_t n = (int64_t)B * C;
    int rc;

    if (c->mode == 1) {
        // forward (+ plan pairs) -> per-bucket sort of both plans -> update of both tables: three launches, one stream
        float* q = reinterpret_cast<float*>(base + c->L.q);
        const DirectPlanDev pi = direct_plan_dev(base + c->L.dir_i[0], n, c->n_items);
        const DirectPlanDev pu = direct_plan_dev(base + c->L.dir_u[0], B, c->n_users);
        profile_begin(B2R_PROF_SCORE_FWD, main_s);
        rc = b2r_bprmf_flash_launch(t->U, uid, t->n_users, t->I, iid, t->n_items, nullptr, g, rows, dQ, q, B, C, d, err_flag,
                                    loss_out, reinterpret_cast<unsigned int*>(base + c->L.counter), &pi, &pu, main_s);
        if (rc == 0) {
            profile_end(B2R_PROF_SCORE_FWD, main_s);
            profile_begin(B2R_PROF_PLAN_I, main_s);
            rc = direct_sort_pair(base + c->L.dir_i[0], n, c->n_items, base + c->L.dir_u[0], B, c->n_users, main_s);
            if (rc != 0) return rc;
            profile_end(B2R_PROF_PLAN_I, main_s);
            const b2r_grad_source su{dQ, nullptr, nullptr, B, 1, 0};
            const b2r_grad_source si{q, g, nullptr, n, C, 0};
            const b2r_apply_job ji{base + c->L.dir_i[0], n, t->n_items, &si, nullptr, nullptr, t->I, t->Im, t->Iv};
            const b2r_apply_job ju{base + c->L.dir_u[0], B, t->n_users, &su, nullptr, nullptr, t->U, t->Um, t->Uv};
            profile_begin(B2R_PROF_SEGMENT_I, main_s);
            rc = direct_apply_pair(&ji, &ju, d, 2, opt, main_s);
            if (rc != 0) return rc;
            profile_end(B2R_PROF_SEGMENT_I, main_s);
            return 0;
        }
        if (rc != B2R_E_UNSUPPORTED) return rc;
        c->mode = 2;                                        // shape outside the streaming kernel's class: legacy form
    }

    // everything the side stream does from here on is ordered after what main has enqueued so far
    // (in particular after the previous step's readers of the plan buffer about to be overwritten)
    B2R_CUDA_OK(cudaEventRecord(c->fork, main_s));
    B2R_CUDA_OK(cudaStreamWaitEvent(c->side, c->fork, 0));

    const int cur = c->slot;
    if (!(c->have_pre && c->pre_uid == uid && c->pre_iid == iid)) {
        rc = build_plans(c, cur, uid, iid, err_flag);     // nothing prefetched for this batch: build it now
        if (rc != 0) return rc;
    }
    // B2R_PLAN_AFTER=1 (A/B knob): start the next batch's plan only when this step's forward kernel has finished, so that
    // the plan's kernels share the SMs with the HBM-bound update kernel instead of with the single-wave forward kernel
    static const bool plan_after = [] { const char* e = getenv("B2R_PLAN_AFTER"); return e && atoi(e) != 0; }();
    const bool prefetch = next_uid != nullptr && next_iid != nullptr;
    if (prefetch && !plan_after) {
        rc = build_plans(c, cur ^ 1, next_uid, next_iid, err_flag);
        if (rc != 0) return rc;
    }
    c->have_pre = prefetch;
    c->pre_uid = prefetch ? next_uid : nullptr;
    c->pre_iid = prefetch ? next_iid : nullptr;
    c->slot = cur ^ 1;

    // main: forward + loss + query-side backward; the gathered user rows are kept in q (the item gradient's source)
    float* q = reinterpret_cast<float*>(base + c->L.q);
    profile_begin(B2R_PROF_SCORE_FWD, main_s);
    rc = b2r_bprmf_fused_fwd_bwd_loss(t->U, uid, t->n_users, t->I, iid, t->n_items, g, rows, dQ, q, B, C, d, err_flag,
                                      loss_out, reinterpret_cast<unsigned int*>(base + c->L.counter), main_s);
    const bool fused = (rc == 0);
    if (rc != 0 && rc != B2R_E_UNSUPPORTED) return rc;
    if (fused) profile_end(B2R_PROF_SCORE_FWD, main_s);
    if (!fused) {
        rc = b2r_gather_rows(t->U, uid, t->n_users, q, B, d, err_flag, main_s);
        if (rc != 0) return rc;
        rc = b2r_rowdot_fwd(q, nullptr, B, t->I, iid, t->n_items, pred, B, C, d, err_flag, main_s);
        if (rc != 0) return rc;
        profile_end(B2R_PROF_SCORE_FWD, main_s);
        profile_begin(B2R_PROF_LOSS, main_s);
        rc = b2r_bpr_loss(pred, loss_out, g, rows, B, C, main_s);
        if (rc != 0) return rc;
        profile_end(B2R_PROF_LOSS, main_s);
        profile_begin(B2R_PROF_SCORE_BWDQ, main_s);
        rc = b2r_rowdot_bwd_query(g, t->I, iid, t->n_items, dQ, B, C, d, main_s);
        if (rc != 0) return rc;
        profile_end(B2R_PROF_SCORE_BWDQ, main_s);
    }

    if (prefetch && plan_after) {
        B2R_CUDA_OK(cudaEventRecord(c->fork, main_s));                 // after the forward kernel(s)
        B2R_CUDA_OK(cudaStreamWaitEvent(c->side, c->fork, 0));
        rc = build_plans(c, cur ^ 1, next_uid, next_iid, err_flag);
        if (rc != 0) return rc;
    }
    // join plan(t); then the fused backward+optimizer on both tables in one launch: dI = g * q reads the saved user
    // rows, so the two updates are independent and the small user-table job runs underneath the item-table job
    B2R_CUDA_OK(cudaStreamWaitEvent(main_s, c->join[cur], 0));
    const PlanBuf& p = c->L.plan[cur];
    const bool dplan = c->mode == 0;
    const b2r_grad_source su{dQ, nullptr, nullptr, B, 1, 0};
    const b2r_grad_source si{q, g, nullptr, n, C, 0};
    const b2r_apply_job ji{base + (dplan ? c->L.dir_i[cur] : p.iws), n, t->n_items, &si, nullptr, nullptr, t->I, t->Im, t->Iv};
    const b2r_apply_job ju{base + (dplan ? c->L.dir_u[cur] : p.uws), B, t->n_users, &su, nullptr, nullptr, t->U, t->Um, t->Uv};
    profile_begin(B2R_PROF_SEGMENT_I, main_s);
    rc = dplan ? direct_apply_pair(&ji, &ju, d, 2, opt, main_s) : b2r_bucket_apply_pair(&ji, &ju, d, 2, opt, main_s);
    if (rc != 0) return rc;
    profile_end(B2R_PROF_SEGMENT_I, main_s);
    return 0;
}
