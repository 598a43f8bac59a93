// bprmf_step.cu -- one whole BPRMF training step enqueued from C (b2r_bprmf_train_step).
//
// Stream plan:   main :  gather u -> rowdot fwd -> loss+grad -> dQ ------------> segment(I) -> segment(U)
//                side :  plan(item ids) -> plan(user ids) ----------------------^ (event join)
// The sort only depends on the ids, so it runs beside the HBM-bound gather kernels on a library-owned
// non-blocking stream; the join is an event wait, never a host synchronisation.
#include "common.cuh"

namespace b2r {

struct StepLayout {
    size_t q, pred, g, rows, dQ;
    size_t ik, ip, is, inu, iws, iws_bytes;
    size_t uk, up, us, unu, uws, uws_bytes;
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
    L->ik = take(n * 4);
    L->ip = take(n * 4);
    L->is = take(n * 4);
    L->inu = take(4);
    L->iws_bytes = b2r_plan_workspace_bytes((int64_t)n, n_items);
    L->iws = take(L->iws_bytes);
    L->uk = take((size_t)B * 4);
    L->up = take((size_t)B * 4);
    L->us = take((size_t)B * 4);
    L->unu = take(4);
    L->uws_bytes = b2r_plan_workspace_bytes(B, n_users);
    L->uws = take(L->uws_bytes);
    L->total = off;
    return L->iws_bytes != 0 && L->uws_bytes != 0;
}

struct SideStream {
    int dev = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};

// one side stream + event pair per device, created on first use and kept for the process lifetime
static int side_stream(SideStream** out) {
    static SideStream cache[16];
    int dev = 0;
    B2R_CUDA_OK(cudaGetDevice(&dev));
    B2R_REQUIRE(dev >= 0 && dev < 16, B2R_E_UNSUPPORTED, "device index %d out of range", dev);
    SideStream& s = cache[dev];
    if (s.dev != dev) {
        B2R_CUDA_OK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        B2R_CUDA_OK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        B2R_CUDA_OK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
        s.dev = dev;
    }
    *out = &s;
    return 0;
}

}  // namespace b2r

using namespace b2r;

extern "C" size_t b2r_bprmf_step_workspace_bytes(int B, int C, int d, int64_t n_users, int64_t n_items) {
    if (B <= 0 || C <= 0 || d <= 0) return 0;
    StepLayout L;
    if (!step_layout(B, C, d, n_users, n_items, &L)) return 0;
    return L.total;
}

extern "C" int b2r_bprmf_train_step(const b2r_bprmf_tables* t, const int64_t* uid, const int64_t* iid, int B,
                                    int C, const b2r_optim* opt, float* loss_out, void* ws, size_t ws_bytes,
                                    int32_t* err_flag, b2r_stream_t stream) {
    B2R_REQUIRE(t && uid && iid && opt && loss_out && ws, B2R_E_BADARG, "b2r_bprmf_train_step: null pointer");
    B2R_REQUIRE(B > 0 && C > 0, B2R_E_BADARG, "b2r_bprmf_train_step: B=%d C=%d", B, C);
    const int d = t->d;
    StepLayout L;
    B2R_REQUIRE(step_layout(B, C, d, t->n_users, t->n_items, &L), B2R_E_UNSUPPORTED,
                "b2r_bprmf_train_step: cannot size the index plans");
    B2R_REQUIRE(ws_bytes >= L.total, B2R_E_WORKSPACE, "b2r_bprmf_train_step: workspace %zu < required %zu",
                ws_bytes, L.total);
    B2R_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, B2R_E_BADARG, "workspace must be 256-byte aligned");
    char* base = static_cast<char*>(ws);
    float* q = reinterpret_cast<float*>(base + L.q);
    float* pred = reinterpret_cast<float*>(base + L.pred);
    float* g = reinterpret_cast<float*>(base + L.g);
    float* rows = reinterpret_cast<float*>(base + L.rows);
    float* dQ = reinterpret_cast<float*>(base + L.dQ);
    uint32_t* ik = reinterpret_cast<uint32_t*>(base + L.ik);
    uint32_t* ip = reinterpret_cast<uint32_t*>(base + L.ip);
    int32_t* is = reinterpret_cast<int32_t*>(base + L.is);
    int32_t* inu = reinterpret_cast<int32_t*>(base + L.inu);
    uint32_t* uk = reinterpret_cast<uint32_t*>(base + L.uk);
    uint32_t* up = reinterpret_cast<uint32_t*>(base + L.up);
    int32_t* us = reinterpret_cast<int32_t*>(base + L.us);
    int32_t* unu = reinterpret_cast<int32_t*>(base + L.unu);

    cudaStream_t main_s = as_stream(stream);
    SideStream* side = nullptr;
    int rc = side_stream(&side);
    if (rc != 0) return rc;
    const int64_t n = (int64_t)B * C;

    // fork: the plans depend only on the ids
    B2R_CUDA_OK(cudaEventRecord(side->fork, main_s));
    B2R_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
    profile_begin(B2R_PROF_PLAN_I, side->stream);
    rc = b2r_plan_build(iid, n, t->n_items, ik, ip, is, inu, base + L.iws, L.iws_bytes, err_flag, side->stream);
    if (rc != 0) return rc;
    profile_end(B2R_PROF_PLAN_I, side->stream);
    rc = b2r_plan_build(uid, B, t->n_users, uk, up, us, unu, base + L.uws, L.uws_bytes, err_flag, side->stream);
    if (rc != 0) return rc;
    B2R_CUDA_OK(cudaEventRecord(side->join, side->stream));

    // main: forward, loss, query-side backward
    rc = b2r_gather_rows(t->U, uid, t->n_users, q, B, d, err_flag, main_s);
    if (rc != 0) return rc;
    profile_begin(B2R_PROF_SCORE_FWD, main_s);
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

    // join, then the fused backward+optimizer on each table (item table first; it only reads the q snapshot)
    B2R_CUDA_OK(cudaStreamWaitEvent(main_s, side->join, 0));
    b2r_grad_source si{q, g, nullptr, n, C, 0 /* ld = d */};
    profile_begin(B2R_PROF_SEGMENT_I, main_s);
    rc = b2r_segment_apply(ik, ip, is, inu, n, t->n_items, d, &si, nullptr, 2, nullptr, nullptr, nullptr, t->I, t->Im, t->Iv,
                           opt, main_s);
    if (rc != 0) return rc;
    profile_end(B2R_PROF_SEGMENT_I, main_s);
    b2r_grad_source su{dQ, nullptr, nullptr, B, 1, 0};
    profile_begin(B2R_PROF_SEGMENT_U, main_s);
    rc = b2r_segment_apply(uk, up, us, unu, B, t->n_users, d, &su, nullptr, 2, nullptr, nullptr, nullptr, t->U, t->Um, t->Uv,
                           opt, main_s);
    if (rc != 0) return rc;
    profile_end(B2R_PROF_SEGMENT_U, main_s);
    return 0;
}
