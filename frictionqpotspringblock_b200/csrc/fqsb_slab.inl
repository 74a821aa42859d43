// fqsb_slab.inl -- slab domain decomposition of ONE very large line / interface over several
// GPUs, inside the library (included by fqsb_api.cu; SURVEY.md section 8e, BASELINE configs #3, #5).
//
// Every member (= one fqsb_system on one GPU) integrates its rows extended by `halo` rows per side
// that mirror the neighbours' rows; a batch of k <= halo steps needs no communication (the local
// array is treated as periodic by the unchanged kernels; the error entering through the outermost
// halo row travels one row per step). After a batch:
//   k_slab_push    writes the member's outermost owned rows (u, v, a, y_l, y_r, index, pcg32
//                  state) STRAIGHT INTO THE NEIGHBOURS' MEMORY over NVLink (peer-mapped stores
//                  into a double-buffered mailbox) and its k x 5 log of per-step sums into every
//                  member's gather mailbox, then publishes an epoch number with st.release.sys;
//   k_slab_import  spins (ld.acquire.sys) on the epochs the neighbours publish, copies the mailbox
//                  into its halo rows, adds the logs of all members in rank order (-> identical
//                  sums, hence identical stop decisions, on every member) into host-mapped memory.
// No NCCL, no host staging, no Python in the exchange: the only host work per batch is the
// StopList replay (detail.h:1764-1784) on the k global sums. The members of a slab live either
// in one process (peer access enabled between the devices) or in one process per GPU (the
// mailboxes are shared through CUDA IPC handles; the caller moves the 64-byte handles).
//
// Per-step decision of the reference (detail.h:1754-1785) batched: the criterion fired at step
// s* < k of a batch -> every member rolls back to its snapshot and redoes exactly s* steps.

#define FQSB_SLAB_MAXW 16
namespace fqsb {

struct SlabDev {
    int rank, world;
    i64 hc; // cells of `halo` rows (one side)
    i64 n;  // local cells (owned + 2 * hc)
    int gcap; // doubles per member in the gather mailbox
    u64* peer[FQSB_SLAB_MAXW]; // mailboxes of all members (peer[rank] = own)
    u64* epoch;           // device [2]: halo / gather epochs completed by this member
    unsigned int* ticket; // device [2]
    double* h_res;        // host-mapped [2 gather parity][world * gcap]
    volatile int* h_status; // host-mapped: [0] peer timeout
    unsigned long long timeout_ns;
};

__host__ __device__ __forceinline__ double* slab_gather(u64* base, int parity, int member, i64 hc,
                                                        int world, int gcap)
{
    return reinterpret_cast<double*>(base + FQSB_SLAB_FLAGS + 28 * hc) +
           (i64)(parity * world + member) * gcap;
}

inline size_t slab_mailbox_words(i64 hc, int world, int gcap)
{
    return (size_t)FQSB_SLAB_FLAGS + 28 * (size_t)hc + 2 * (size_t)world * (size_t)gcap;
}

// spin until *flag >= epoch (published by a peer with st.release.sys); gives up after timeout_ns
__device__ __forceinline__ bool slab_wait(const u64* flag, u64 epoch, const SlabDev& D)
{
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < epoch) {
        if (global_ns() - t0 > D.timeout_ns) {
            D.h_status[0] = 1;
            __threadfence_system();
            return false;
        }
        __nanosleep(200);
    }
    return true;
}


// 7 planes x 2 sides of halo cells in one pass: every thread issues its 14 loads before its 14
// stores (16-byte accesses when everything is 16-byte aligned), so a member's ~15 MB of halo rows
// (4096^2 interface, 32 rows) move at NVLink / HBM rate instead of at the latency of one
// load-store pair per thread at a time.
template <bool LDCG, class SA, class DA, class SB, class DB>
__device__ __forceinline__ void slab_copy14(SA srcA, DA dstA, SB srcB, DB dstB, const i64 cells,
                                            const bool vec)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 t0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        for (i64 c = t0; c < (cells >> 1); c += stride) {
            ulonglong2 x[14];
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const ulonglong2* a = reinterpret_cast<const ulonglong2*>(srcA(q)) + c;
                const ulonglong2* b = reinterpret_cast<const ulonglong2*>(srcB(q)) + c;
                x[q] = LDCG ? __ldcg(a) : *a;
                x[7 + q] = LDCG ? __ldcg(b) : *b;
            }
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                reinterpret_cast<ulonglong2*>(dstA(q))[c] = x[q];
                reinterpret_cast<ulonglong2*>(dstB(q))[c] = x[7 + q];
            }
        }
    }
    else {
        for (i64 c = t0; c < cells; c += stride) {
            u64 x[14];
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                x[q] = LDCG ? __ldcg(srcA(q) + c) : srcA(q)[c];
                x[7 + q] = LDCG ? __ldcg(srcB(q) + c) : srcB(q)[c];
            }
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                dstA(q)[c] = x[q];
                dstB(q)[c] = x[7 + q];
            }
        }
    }
}

// ---- after a batch: outermost owned rows -> the neighbours' mailboxes, log -> all mailboxes ------
__global__ void __launch_bounds__(256)
    k_slab_push(const State S, const SlabDev D, const double* log, int nlog, int halos)
{
    __shared__ int s_last;
    const u64 e = D.epoch[0] + 1;
    const int par = (int)(e & 1ULL);
    const int prev = (D.rank + D.world - 1) % D.world, next = (D.rank + 1) % D.world;
    // my top rows become the previous member's bottom halo (its side 1 = "from next"), my bottom
    // rows the next member's top halo (side 0 = "from prev")
    u64* to_prev = slab_mail(D.peer[prev], par, 1, D.hc);
    u64* to_next = slab_mail(D.peer[next], par, 0, D.hc);
    const u64* plane[7] = {reinterpret_cast<const u64*>(S.u),  reinterpret_cast<const u64*>(S.v),
                           reinterpret_cast<const u64*>(S.a),  reinterpret_cast<const u64*>(S.yl),
                           reinterpret_cast<const u64*>(S.yr), reinterpret_cast<const u64*>(S.idx),
                           S.rng};
    const i64 top = D.hc, bot = D.n - 2 * D.hc;
    if (halos) {
        const i64 hc = D.hc;
        slab_copy14<false>([&](int q) { return plane[q] + top; },
                           [&](int q) { return to_prev + (i64)q * hc; },
                           [&](int q) { return plane[q] + bot; },
                           [&](int q) { return to_next + (i64)q * hc; }, hc,
                           ((D.hc | D.n) & 1) == 0);
    }
    const u64 ge = D.epoch[1] + 1;
    if (nlog > 0 && blockIdx.x == 0) {
        const int gpar = (int)(ge & 1ULL);
        for (int j = 0; j < D.world; ++j) {
            double* dst = slab_gather(D.peer[j], gpar, D.rank, D.hc, D.world, D.gcap);
            for (int i = threadIdx.x; i < nlog; i += blockDim.x) {
                dst[i] = log[i];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(&D.ticket[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) {
        return;
    }
    // the last CTA publishes the epochs: everything above is visible system-wide before them
    __threadfence_system();
    if (threadIdx.x == 0) {
        D.ticket[0] = 0u;
        if (halos) {
            st_release_sys(D.peer[prev] + 1, e); // prev's flag "from next"
            st_release_sys(D.peer[next] + 0, e); // next's flag "from prev"
        }
    }
    if (nlog > 0 && threadIdx.x < D.world) {
        st_release_sys(D.peer[threadIdx.x] + 2 + D.rank, ge);
    }
}

// ---- before the next batch: mailbox -> halo rows; logs of all members -> host-mapped result ------
// raw == 0: h_res[i] = sum over members (rank order) of their entry i; raw != 0: h_res[j*nlog + i]
__global__ void __launch_bounds__(256)
    k_slab_import(const State S, const SlabDev D, int nlog, int raw, int halos)
{
    __shared__ int s_ok, s_last;
    const u64 e = D.epoch[0] + 1;
    const u64 ge = D.epoch[1] + 1;
    u64* self = D.peer[D.rank];
    if (halos) {
        const int par = (int)(e & 1ULL);
        if (threadIdx.x == 0) {
            s_ok = slab_wait(self + 0, e, D) && slab_wait(self + 1, e, D);
        }
        __syncthreads();
        if (s_ok) {
            const u64* from_prev = slab_mail(self, par, 0, D.hc);
            const u64* from_next = slab_mail(self, par, 1, D.hc);
            u64* plane[7] = {reinterpret_cast<u64*>(S.u),  reinterpret_cast<u64*>(S.v),
                             reinterpret_cast<u64*>(S.a),  reinterpret_cast<u64*>(S.yl),
                             reinterpret_cast<u64*>(S.yr), reinterpret_cast<u64*>(S.idx), S.rng};
            const i64 bot = D.n - D.hc, hc = D.hc;
            slab_copy14<true>([&](int q) { return from_prev + (i64)q * hc; },
                              [&](int q) { return plane[q]; },
                              [&](int q) { return from_next + (i64)q * hc; },
                              [&](int q) { return plane[q] + bot; }, hc, ((D.hc | D.n) & 1) == 0);
        }
    }
    if (nlog > 0 && blockIdx.x == 0) {
        const int gpar = (int)(ge & 1ULL);
        double* res = D.h_res + (size_t)gpar * D.world * D.gcap;
        __syncthreads();
        if (threadIdx.x == 0) {
            s_ok = 1;
        }
        __syncthreads();
        if (threadIdx.x < D.world) {
            if (!slab_wait(self + 2 + threadIdx.x, ge, D)) {
                s_ok = 0;
            }
        }
        __syncthreads();
        if (s_ok) {
            for (int i = threadIdx.x; i < nlog; i += blockDim.x) {
                if (raw) {
                    for (int j = 0; j < D.world; ++j) {
                        res[j * nlog + i] =
                            __ldcg(slab_gather(self, gpar, j, D.hc, D.world, D.gcap) + i);
                    }
                }
                else {
                    double acc = 0.0;
                    for (int j = 0; j < D.world; ++j) {
                        acc += __ldcg(slab_gather(self, gpar, j, D.hc, D.world, D.gcap) + i);
                    }
                    res[i] = acc;
                }
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(&D.ticket[1], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        D.ticket[1] = 0u;
        if (halos) {
            D.epoch[0] = e;
        }
        if (nlog > 0) {
            D.epoch[1] = ge;
        }
    }
}

// =============================================================================================
// 1-D lines: every member runs the temporally blocked kernel (fqsb_blocked.cuh) on its local line,
// ONE launch per batch of k steps. That kernel writes the complete new state into the other buffer
// set, so a batch is revocable without a snapshot, and the commit / redo protocol of the single-GPU
// kernel carries over: the stop decision moves from k_blocked's last tile into k_slab_import_blocked
// -- after the members' per-step sums have been gathered -- where every member replays
// timeStepsUntilEvent / minimise (detail.h:1605-1619, 1764-1784) on identical global sums and
// either commits the batch (flips the buffer set) or asks for a batch of exactly s* steps. The host
// only enqueues batches and polls the status.
// =============================================================================================
struct SlabPlanes {
    u64* p[2][7]; // the two buffer sets: u, v, a, y_l, y_r, idx, rng
};

// set_arg 0 / 1: that buffer set; 2: the set the logged batch just wrote (Ctl::flip ^ 1)
__global__ void __launch_bounds__(256)
    k_slab_push_blocked(const State S, const SlabDev D, const SlabPlanes PL, const int set_arg,
                        const double* log, const int logged)
{
    __shared__ int s_last;
    const int set = set_arg < 2 ? set_arg : (S.ctl[0].flip ^ 1);
    const int nlog = logged ? S.ctl[0].batch * FQSB_NLOG : 0;
    const u64 e = D.epoch[0] + 1;
    const int par = (int)(e & 1ULL);
    const int prev = (D.rank + D.world - 1) % D.world, next = (D.rank + 1) % D.world;
    u64* to_prev = slab_mail(D.peer[prev], par, 1, D.hc);
    u64* to_next = slab_mail(D.peer[next], par, 0, D.hc);
    const i64 top = D.hc, bot = D.n - 2 * D.hc;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 t0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        const u64* src = PL.p[set][q];
        for (i64 c = t0; c < D.hc; c += stride) {
            to_prev[(i64)q * D.hc + c] = src[top + c];
            to_next[(i64)q * D.hc + c] = src[bot + c];
        }
    }
    const u64 ge = D.epoch[1] + 1;
    if (nlog > 0 && blockIdx.x == 0) {
        const int gpar = (int)(ge & 1ULL);
        for (int j = 0; j < D.world; ++j) {
            double* dst = slab_gather(D.peer[j], gpar, D.rank, D.hc, D.world, D.gcap);
            for (int i = threadIdx.x; i < nlog; i += blockDim.x) {
                dst[i] = log[i];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(&D.ticket[0], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) {
        return;
    }
    __threadfence_system();
    if (threadIdx.x == 0) {
        D.ticket[0] = 0u;
        st_release_sys(D.peer[prev] + 1, e);
        st_release_sys(D.peer[next] + 0, e);
    }
    if (logged && threadIdx.x < D.world) { // (published even when nlog == 0: the epochs stay in step)
        st_release_sys(D.peer[threadIdx.x] + 2 + D.rank, ge);
    }
}

__global__ void __launch_bounds__(256)
    k_slab_import_blocked(const State S, const SlabDev D, const SlabPlanes PL, const int set_arg,
                          const int logged, const RunArgs A, const int ksteps,
                          const int advance = 1)
{
    __shared__ int s_ok, s_last;
    __shared__ double tot[FQSB_BK_MAXSTEPS * FQSB_NLOG];
    Ctl& ctl = S.ctl[0];
    const int flip = ctl.flip;
    const int set = set_arg < 2 ? set_arg : (flip ^ 1);
    const int nsteps = logged ? ctl.batch : 0;
    const bool running = ctl.status == ST_RUNNING;
    // (advance > 1: the halo exchanges in between rode on the tile kernel, BlockedFuse)
    const u64 e = D.epoch[0] + (u64)advance;
    const u64 ge = D.epoch[1] + 1;
    u64* self = D.peer[D.rank];
    const int par = (int)(e & 1ULL);
    if (threadIdx.x == 0) {
        s_ok = slab_wait(self + 0, e, D) && slab_wait(self + 1, e, D);
    }
    __syncthreads();
    if (s_ok) {
        const u64* from_prev = slab_mail(self, par, 0, D.hc);
        const u64* from_next = slab_mail(self, par, 1, D.hc);
        const i64 stride = (i64)gridDim.x * blockDim.x;
        const i64 t0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
        const i64 bot = D.n - D.hc;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            u64* dst = PL.p[set][q];
            for (i64 c = t0; c < D.hc; c += stride) {
                dst[c] = __ldcg(from_prev + (i64)q * D.hc + c);
                dst[bot + c] = __ldcg(from_next + (i64)q * D.hc + c);
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(&D.ticket[1], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) {
        return;
    }
    // ---- the last CTA: every halo row of this member is in place. Gather the members' sums
    //      (rank order), replay the per-step decisions, commit or ask for a redo
    __threadfence();
    if (logged) {
        const int gpar = (int)(ge & 1ULL);
        if (threadIdx.x == 0) {
            s_ok = 1;
        }
        __syncthreads();
        if (threadIdx.x < D.world) {
            if (!slab_wait(self + 2 + threadIdx.x, ge, D)) {
                s_ok = 0;
            }
        }
        __syncthreads();
        const int nlog = nsteps * FQSB_NLOG;
        for (int i = threadIdx.x; i < nlog; i += blockDim.x) {
            double acc = 0.0;
            for (int j = 0; j < D.world; ++j) {
                acc += __ldcg(slab_gather(self, gpar, j, D.hc, D.world, D.gcap) + i);
            }
            tot[i] = acc;
        }
        __syncthreads();
        if (threadIdx.x < 32 && running && s_ok && nsteps > 0) {
            const int lane = threadIdx.x;
            Prog g;
            prog_load(g, ctl);
            RingEntry ring = ring_load(ctl, A, lane);
            int status = ST_RUNNING;
            double sf = 0.0, sff = 0.0;
            int s = 0;
            for (; s < nsteps; ++s) {
                sf = tot[s * FQSB_NLOG];
                sff = tot[s * FQSB_NLOG + 1];
                g.inc++; // detail.h:1541
                status = step_decide(A, g, ring, lane, sf, sff, (int)tot[s * FQSB_NLOG + 2],
                                     (int)tot[s * FQSB_NLOG + 3], (int)tot[s * FQSB_NLOG + 4]);
                if (status != ST_RUNNING) {
                    break;
                }
            }
            if (status == ST_RUNNING || s == nsteps - 1) { // commit
                ring_store(ctl, A, lane, ring);
                if (lane == 0) {
                    prog_store(g, ctl);
                    ctl.residual = residual_from_sums(sf, sff);
                    ctl.flip = flip ^ 1;
                    const i64 left = A.max_steps - g.steps;
                    ctl.batch = (int)(left < ksteps ? left : ksteps);
                    ctl.status = status;
                }
            }
            else if (lane == 0) { // the criterion fired inside the batch: redo exactly s + 1 steps
                ctl.batch = s + 1;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        D.ticket[1] = 0u;
        D.epoch[0] = e;
        if (logged) {
            D.epoch[1] = ge;
            D.h_status[1] = ctl.status; // the host stops enqueueing batches once this leaves RUNNING
        }
    }
}

// a logged batch leaves per-CTA partials [k][tiles][FQSB_NPART] (stream_finalise / the sweep
// kernels in slot mode): add them up per step (CTAs in a fixed order) -> log [k][FQSB_NLOG], and
// settle the bookkeeping the per-step finalise would have kept
__global__ void __launch_bounds__(256)
    k_slab_reduce_log(const State S, const double* part, int tiles, double* log, int k,
                      int overdamped)
{
    __shared__ double scratch[32 * FQSB_NLOG];
    const int j = blockIdx.x;
    double tot[FQSB_NLOG] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const double* e = part + (size_t)j * tiles * FQSB_NPART;
    for (int c = threadIdx.x; c < tiles; c += blockDim.x) {
#pragma unroll
        for (int q = 0; q < FQSB_NLOG; ++q) {
            tot[q] += e[(size_t)c * FQSB_NPART + q];
        }
    }
    block_sum<FQSB_NLOG>(tot, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < FQSB_NLOG; ++q) {
            log[j * FQSB_NLOG + q] = tot[q];
        }
        if (j == 0) {
            Ctl& c = S.ctl[0];
            if (!overdamped) {
                c.inc += k; // detail.h:1541
            }
            c.steps = k;
            c.flip = k & 1;
            c.count = 0u;
            c.status = ST_EXHAUSTED;
        }
    }
}

// snapshot / rollback of the whole local state in one launch (dir 0: state -> snapshot)
struct SnapArgs {
    u64* a[7];
    u64* b[7];
    double *uf_a, *uf_b;
    Ctl *ctl_a, *ctl_b;
    i64 n;
};

__global__ void __launch_bounds__(256) k_slab_copy_state(const SnapArgs X)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 t0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if ((X.n & 1) == 0) {
        const i64 n2 = X.n >> 1;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(X.a[q]);
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(X.b[q]);
            for (i64 c = t0; c < n2; c += stride) {
                dst[c] = src[c];
            }
        }
    }
    else {
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            for (i64 c = t0; c < X.n; c += stride) {
                X.b[q][c] = X.a[q][c];
            }
        }
    }
    if (t0 == 0) {
        *X.uf_b = *X.uf_a;
        *X.ctl_b = *X.ctl_a;
    }
}

// u_frame += duf; du[0] = dup (the uniform shift k_align applies)
__global__ void k_slab_shift(const State S, double* du, double dup, double duf)
{
    du[0] = dup;
    S.u_frame[0] += duf;
}

} // namespace fqsb

struct SlabSnap {
    u64* p[7];
    double* uf;
    Ctl* ctl;
};

struct SlabGraph { // CUDA graph of one batch shape (snapshot + k steps + reduce + push)
    i64 k;
    int gmode;
    bool seen;
    cudaGraphExec_t exec;
};
#define FQSB_SLAB_GRAPHS 6

struct fqsb_slab_state {
    int rank, world;
    i64 halo_cells;
    int kmax, gcap;
    u64* mailbox;
    size_t mailbox_bytes;
    u64* peer[FQSB_SLAB_MAXW];
    bool peer_ipc[FQSB_SLAB_MAXW];
    bool connected;
    u64* d_epoch;
    unsigned int* d_ticket;
    double* h_res;
    int* h_status;
    SlabDev dev;
    bool overdamped;
    bool blocked;       // 1-D line: members run the temporally blocked kernel
    SlabSnap snap[2];   // two snapshots: a speculative batch keeps its predecessor's intact
    double* d_part;     // [kmax][tiles][FQSB_NPART] per-CTA partials of a logged batch
    int part_tiles;
    cudaEvent_t ev[2];  // recorded after the import of a batch
    cudaEvent_t evb[4]; // blocked path: ring of events bounding the host's run-ahead
    u64 gathers;        // gather epochs enqueued so far (host mirror of d_epoch[1])
    SlabGraph graphs[FQSB_SLAB_GRAPHS];
    i64 batches, redone, wasted;
};

static int slab_flags_all_fwd(fqsb_system** m, int nm);

static int slab_require(fqsb_system* s, bool connected = true)
{
    TRY(enter(s));
    if (!s->slab) {
        return fail(FQSB_EASSERT, "not a slab member (call fqsb_slab_init first)");
    }
    if (connected && !s->slab->connected) {
        return fail(FQSB_EASSERT, "slab members are not connected (fqsb_slab_connect)");
    }
    return FQSB_OK;
}

static void slab_free(fqsb_system* s)
{
    fqsb_slab_state* L = s->slab;
    if (!L) {
        return;
    }
    for (int j = 0; j < L->world; ++j) {
        if (L->peer_ipc[j] && L->peer[j]) {
            cudaIpcCloseMemHandle(L->peer[j]);
        }
    }
    for (int k = 0; k < FQSB_SLAB_GRAPHS; ++k) {
        if (L->graphs[k].exec) {
            cudaGraphExecDestroy(L->graphs[k].exec);
        }
    }
    for (int k = 0; k < 2; ++k) {
        if (L->ev[k]) {
            cudaEventDestroy(L->ev[k]);
        }
    }
    for (int k = 0; k < 4; ++k) {
        if (L->evb[k]) {
            cudaEventDestroy(L->evb[k]);
        }
    }
    if (L->mailbox) {
        cudaFree(L->mailbox);
    }
    if (L->d_epoch) {
        cudaFree(L->d_epoch);
    }
    if (L->d_ticket) {
        cudaFree(L->d_ticket);
    }
    if (L->h_res) {
        cudaFreeHost(L->h_res);
    }
    if (L->h_status) {
        cudaFreeHost(L->h_status);
    }
    delete L;
    s->slab = nullptr;
}

// all launches of `k` steps (Verlet, or no-passing sweeps) on the streaming kernels, without any
// host synchronisation. MODE_LOG: every launch leaves its per-CTA partial sums over the owned range
// in its own slot, k_slab_reduce_log adds them up -> s->d_log [k][FQSB_NLOG]; MODE_FIXED: plain steps.
static int slab_enqueue_steps(fqsb_system* s, i64 k, int mode, int flow, double v_frame)
{
    fqsb_slab_state* L = s->slab;
    const bool overdamped = s->par.minimisation == FQSB_MIN_OVERDAMPED;
    RunArgs A = make_args(mode, k);
    A.flow = flow;
    A.v_frame = v_frame;
    A.own_lo = (int)s->own_lo;
    A.own_hi = (int)s->own_hi;
    A.log = mode == MODE_LOG ? L->d_part : nullptr;
    const bool slots = mode == MODE_LOG;
    const unsigned rg = (unsigned)((s->R + 127) / 128);
    k_ctl_begin<<<rg, 128, 0, s->stream>>>(s->P, s->S, 0, overdamped ? 1 : 0, s->d_out);
    const int finalise = flow ? 1 : 0;
    const i64 nl = overdamped ? k + 1 : k;
    for (i64 b = 0; b < nl; ++b) {
        cudaError_t e;
        if (overdamped) {
            const int sweep = (b < k ? 1 : 0) | (slots ? (int)((b > 0 ? b : 1) << 1) : 0);
            e = launch_stream_sweep(s->P, s->S, A, s->stream, (int)(b & 1), b == 0, sweep);
        }
        else {
            e = launch_stream_step(s->P, s->S, A, s->stream, (int)(b & 1),
                                   slots ? (2 | (int)(b << 2)) : finalise);
        }
        if (e != cudaSuccess) {
            return cuda_fail(e, "stream kernel launch");
        }
    }
    if (slots) {
        k_slab_reduce_log<<<(unsigned)k, 256, 0, s->stream>>>(s->S, L->d_part, L->part_tiles,
                                                              s->d_log, (int)k, overdamped ? 1 : 0);
    }
    else if (!finalise && !overdamped) {
        k_stream_fixed_done<<<rg, 128, 0, s->stream>>>(s->P, s->S, k, 0);
    }
    dim3 grid((unsigned)s->S.tiles, (unsigned)s->R);
    k_stream_settle<<<grid, 256, 0, s->stream>>>(s->P, s->S, overdamped ? 0 : 1);
    k_stream_settle_flags<<<rg, 128, 0, s->stream>>>(s->P, s->S);
    CU(cudaGetLastError());
    s->launches += nl + 4;
    s->steps += k;
    s->last_kernel = overdamped ? "slab_stream_nopassing" : "slab_stream";
    invalidate_forces(s);
    return FQSB_OK;
}

// slot: which of the two snapshots; restore = snapshot -> live state
static int slab_copy_state(fqsb_system* s, int slot, bool restore)
{
    const SlabSnap& Z = s->slab->snap[slot];
    SnapArgs X;
    u64* live[7] = {(u64*)s->S.u, (u64*)s->S.v, (u64*)s->S.a, (u64*)s->S.yl, (u64*)s->S.yr,
                    (u64*)s->S.idx, s->S.rng};
    for (int q = 0; q < 7; ++q) {
        X.a[q] = restore ? Z.p[q] : live[q];
        X.b[q] = restore ? live[q] : Z.p[q];
    }
    X.uf_a = restore ? Z.uf : s->S.u_frame;
    X.uf_b = restore ? s->S.u_frame : Z.uf;
    X.ctl_a = restore ? Z.ctl : s->S.ctl;
    X.ctl_b = restore ? s->S.ctl : Z.ctl;
    X.n = s->n;
    k_slab_copy_state<<<148 * 4, 256, 0, s->stream>>>(X);
    CU(cudaGetLastError());
    s->launches++;
    if (restore) {
        invalidate_forces(s);
    }
    return FQSB_OK;
}

static int slab_enqueue_push(fqsb_system* s, int nlog, const double* src, int halos)
{
    fqsb_slab_state* L = s->slab;
    // one 16-byte word of each of the 14 (plane, side) pairs per thread
    unsigned grid = halos ? (unsigned)((L->halo_cells / 2 + 255) / 256) : 1u;
    grid = grid < 1u ? 1u : (grid > 592u ? 592u : grid);
    k_slab_push<<<grid, 256, 0, s->stream>>>(s->S, L->dev, src, nlog, halos);
    CU(cudaGetLastError());
    s->launches++;
    return FQSB_OK;
}

// returns (through *res) where the gathered values of this import will appear on the host
static int slab_enqueue_import(fqsb_system* s, int nlog, int raw, int halos,
                               const double** res = nullptr, int ev = -1)
{
    fqsb_slab_state* L = s->slab;
    // (every CTA spins on the neighbours' flags -- set by OTHER devices, so the CTAs need not be
    // co-resident -- then moves one 16-byte word of each (plane, side) pair per thread)
    unsigned grid = halos ? (unsigned)((L->halo_cells / 2 + 255) / 256) : 1u;
    grid = grid < 1u ? 1u : (grid > 592u ? 592u : grid);
    k_slab_import<<<grid, 256, 0, s->stream>>>(s->S, L->dev, nlog, raw, halos);
    CU(cudaGetLastError());
    s->launches++;
    if (nlog > 0) {
        L->gathers++;
        if (res) {
            *res = L->h_res + (size_t)(L->gathers & 1ULL) * L->world * L->gcap;
        }
    }
    if (ev >= 0) {
        CU(cudaEventRecord(L->ev[ev], s->stream));
    }
    if (halos) {
        invalidate_forces(s);
    }
    return FQSB_OK;
}

static int slab_status(fqsb_system* s)
{
    if (s->slab->h_status[0]) {
        s->slab->h_status[0] = 0;
        return fail(FQSB_ECUDA, "slab: timed out waiting for a neighbouring member");
    }
    return FQSB_OK;
}

static int slab_sync(fqsb_system* s)
{
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    return slab_status(s);
}

static int slab_wait_event(fqsb_system* s, int ev, bool blocked_ring = false)
{
    CU(cudaSetDevice(s->device));
    CU(cudaEventSynchronize(blocked_ring ? s->slab->evb[ev] : s->slab->ev[ev]));
    return slab_status(s);
}

// one batch on one member: [snapshot into `snap_slot`] + k steps + push, as direct launches or as
// the replay of a CUDA graph captured on the second batch of the same shape
static int slab_enqueue_batch(fqsb_system* s, i64 k, int mode, int snap_slot, int flow,
                              double v_frame)
{
    CU(cudaSetDevice(s->device));
    fqsb_slab_state* L = s->slab;
    const int nlog = mode == MODE_LOG ? (int)(k * FQSB_NLOG) : 0;
    const int gmode = mode * 8 + (snap_slot + 1) * 2 + (flow ? 1 : 0);
    static const bool use_graph = [] {
        const char* e = std::getenv("FQSB_SLAB_GRAPH");
        return e ? std::atoi(e) != 0 : true;
    }();
    auto body = [&]() -> int {
        if (snap_slot >= 0) {
            TRY(slab_copy_state(s, snap_slot, false));
        }
        TRY(slab_enqueue_steps(s, k, mode, flow, v_frame));
        TRY(slab_enqueue_push(s, nlog, s->d_log, 1));
        return FQSB_OK;
    };
    if (!use_graph || flow) { // (flowSteps: v_frame is a kernel argument, not worth a graph)
        return body();
    }
    SlabGraph* G = nullptr;
    for (int i = 0; i < FQSB_SLAB_GRAPHS; ++i) {
        if (L->graphs[i].seen && L->graphs[i].k == k && L->graphs[i].gmode == gmode) {
            G = &L->graphs[i];
        }
    }
    if (!G) {
        // first batch of this shape: direct launches (allocations, function attributes); shapes
        // seen once only (e.g. the redone tail of a minimisation) are recycled first
        for (int i = 0; i < FQSB_SLAB_GRAPHS && !G; ++i) {
            if (!L->graphs[i].seen) {
                G = &L->graphs[i];
            }
        }
        for (int i = 0; i < FQSB_SLAB_GRAPHS && !G; ++i) {
            if (!L->graphs[i].exec) {
                G = &L->graphs[i];
            }
        }
        if (!G) {
            G = &L->graphs[0];
            cudaGraphExecDestroy(G->exec);
        }
        G->exec = nullptr;
        G->seen = true;
        G->k = k;
        G->gmode = gmode;
        TRY(body());
        // ... and captured right away (the GPU is busy with the batch just enqueued), so that the
        // next batch of this shape is a single graph launch
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        const i64 launches0 = s->launches, steps0 = s->steps;
        int rc = body();
        cudaError_t ce = cudaStreamEndCapture(s->stream, &g);
        s->launches = launches0;
        s->steps = steps0;
        if (rc == FQSB_OK && ce == cudaSuccess && g &&
            cudaGraphInstantiate(&G->exec, g, 0) != cudaSuccess) {
            G->exec = nullptr;
        }
        if (g) {
            cudaGraphDestroy(g);
        }
        cudaGetLastError();
        return FQSB_OK;
    }
    if (!G->exec) { // (the capture after the first batch failed: try once more)
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        const i64 launches0 = s->launches, steps0 = s->steps;
        int rc = body();
        cudaError_t ce = cudaStreamEndCapture(s->stream, &g);
        s->launches = launches0;
        s->steps = steps0;
        if (rc != FQSB_OK || ce != cudaSuccess || !g) {
            if (g) {
                cudaGraphDestroy(g);
            }
            cudaGetLastError();
            G->seen = false;
            return rc != FQSB_OK ? rc : cuda_fail(ce, "slab graph capture");
        }
        ce = cudaGraphInstantiate(&G->exec, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) {
            G->exec = nullptr;
            G->seen = false;
            return cuda_fail(ce, "cudaGraphInstantiate");
        }
    }
    CU(cudaGraphLaunch(G->exec, s->stream));
    s->launches += k + 7;
    s->steps += k;
    invalidate_forces(s);
    return FQSB_OK;
}

// ---- 1-D lines on the temporally blocked kernel ------------------------------------------------
static SlabPlanes slab_planes(const fqsb_system* s)
{
    SlabPlanes PL;
    u64* a[7] = {(u64*)s->S.u, (u64*)s->S.v, (u64*)s->S.a, (u64*)s->S.yl, (u64*)s->S.yr,
                 (u64*)s->S.idx, s->S.rng};
    u64* b[7] = {(u64*)s->S.u2, (u64*)s->S.v2, (u64*)s->S.a2, (u64*)s->bk.yl2, (u64*)s->bk.yr2,
                 (u64*)s->bk.idx2, s->bk.rng2};
    for (int q = 0; q < 7; ++q) {
        PL.p[0][q] = a[q];
        PL.p[1][q] = b[q];
    }
    return PL;
}

static unsigned slab_copy_grid(const fqsb_slab_state* L, unsigned cap)
{
    const i64 words = 14 * L->halo_cells;
    unsigned grid = (unsigned)((words + 4095) / 4096);
    return grid < 1u ? 1u : (grid > cap ? cap : grid);
}

// One dynamics call of a slab of blocked members. A.mode == MODE_FIXED: n = A.max_steps steps in
// batches of `batch`; stop modes: batches until every member's (identical) status leaves
// ST_RUNNING. On return the current state sits in the primary buffer set and h_ctl is up to date.
static int slab_blocked_run(fqsb_system** m, int nm, RunArgs A, i64 batch)
{
    if (batch > FQSB_BK_MAXSTEPS) {
        batch = FQSB_BK_MAXSTEPS;
    }
    std::vector<BlockedPlan> plan((size_t)nm);
    for (int g = 0; g < nm; ++g) {
        fqsb_system* s = m[g];
        CU(cudaSetDevice(s->device));
        plan[(size_t)g] = blocked_plan(s->P, (int)batch, (s->par.kernel >> 16) & 0xffff);
        const BlockedPlan& pl = plan[(size_t)g];
        if (pl.B < 2 || pl.B > 8 || blocked_smem(pl.B) > 227 * 1024 || pl.ksteps < batch) {
            return fail(FQSB_EUNSUPPORTED, "no tile geometry for the blocked kernel");
        }
        TRY(ensure_blocked_buffers(s, pl));
        const unsigned rg = (unsigned)((s->R + 127) / 128);
        k_ctl_begin<<<rg, 128, 0, s->stream>>>(s->P, s->S, 0, 0, s->d_out);
        CU(cudaGetLastError());
        s->launches++;
        s->last_kernel = "slab_blocked_1d";
        invalidate_forces(s);
    }
    A.own_lo = (int)m[0]->own_lo;
    A.own_hi = (int)m[0]->own_hi;
    if (A.mode == MODE_FIXED) {
        // Fixed steps need no decision between batches, so the halo exchange rides on the tile
        // kernel (BlockedFuse): the tiles at the member's ends read the halo regions from the
        // mailbox and write the member's outermost owned blocks into the neighbours' mailboxes
        // from their write-back. ONE launch per batch; one import at the end of the call brings
        // the last mail into the state arrays. FQSB_SLAB_FUSE=0: push / import launches instead.
        static const bool fuse = [] {
            const char* e = getenv("FQSB_SLAB_FUSE");
            return !(e && e[0] == '0');
        }();
        int parity = 0;
        i64 left = A.max_steps;
        int nb = 0;
        while (left > 0) {
            const i64 k = left < batch ? left : batch;
            for (int g = 0; g < nm; ++g) {
                fqsb_system* s = m[g];
                CU(cudaSetDevice(s->device));
                RunArgs Ag = A;
                Ag.own_lo = (int)s->own_lo;
                Ag.own_hi = (int)s->own_hi;
                BlockedArgs K = s->bk;
                K.nsteps = (int)k;
                K.flip = parity;
                if (fuse) {
                    const SlabDev& D = s->slab->dev;
                    const BlockedPlan& pl = plan[(size_t)g];
                    BlockedFuse& F = K.fuse;
                    F.on = 1;
                    F.pull = nb > 0;
                    F.batch = nb;
                    F.hc = D.hc;
                    F.self = D.peer[D.rank];
                    F.prev = D.peer[(D.rank + D.world - 1) % D.world];
                    F.next = D.peer[(D.rank + 1) % D.world];
                    F.epoch = D.epoch;
                    F.count = D.ticket + 2;
                    F.h_status = D.h_status;
                    F.timeout_ns = D.timeout_ns;
                    F.n_readers = F.n_pushers = 0;
                    for (int c = 0; c < pl.ntiles; ++c) {
                        bool rd, pu;
                        blocked_tile_roles(s->P.N, D.hc, pl.own, pl.H, c, &rd, &pu);
                        F.n_readers += rd && F.pull;
                        F.n_pushers += pu;
                    }
                }
                cudaError_t e = launch_blocked(plan[(size_t)g], s->P, s->S, Ag, K, s->stream);
                if (e != cudaSuccess) {
                    return cuda_fail(e, "blocked kernel launch");
                }
                s->launches++;
                if (!fuse) {
                    k_slab_push_blocked<<<slab_copy_grid(s->slab, 64u), 256, 0, s->stream>>>(
                        s->S, s->slab->dev, slab_planes(s), parity ^ 1, nullptr, 0);
                    CU(cudaGetLastError());
                    s->launches++;
                }
                s->steps += k;
            }
            if (!fuse) {
                for (int g = 0; g < nm; ++g) {
                    fqsb_system* s = m[g];
                    CU(cudaSetDevice(s->device));
                    k_slab_import_blocked<<<slab_copy_grid(s->slab, 32u), 256, 0, s->stream>>>(
                        s->S, s->slab->dev, slab_planes(s), parity ^ 1, 0, A, (int)batch);
                    CU(cudaGetLastError());
                    s->launches++;
                }
            }
            parity ^= 1;
            left -= k;
            nb++;
        }
        if (fuse && nb > 0) {
            for (int g = 0; g < nm; ++g) {
                fqsb_system* s = m[g];
                CU(cudaSetDevice(s->device));
                k_slab_import_blocked<<<slab_copy_grid(s->slab, 32u), 256, 0, s->stream>>>(
                    s->S, s->slab->dev, slab_planes(s), parity, 0, A, (int)batch, nb);
                CU(cudaGetLastError());
                s->launches++;
            }
        }
        for (int g = 0; g < nm; ++g) {
            fqsb_system* s = m[g];
            CU(cudaSetDevice(s->device));
            CU(launch_blocked_fixed_done(s->P, s->S, A.max_steps, parity, s->stream));
            s->launches++;
        }
    }
    else {
        RunArgs Alog = A;
        Alog.mode = MODE_LOG;
        Alog.track = 0;
        for (int g = 0; g < nm; ++g) {
            fqsb_system* s = m[g];
            CU(cudaSetDevice(s->device));
            CU(launch_blocked_begin(s->P, s->S, (int)batch, A.max_steps, s->stream));
            s->launches++;
        }
        // The host runs at most 3 batches ahead of the device: batch b is enqueued once batch b - 3
        // has finished, and only while the status the import kernel mirrors into host-mapped
        // memory says "running". No synchronisation of the whole stream inside the loop.
        for (int g = 0; g < nm; ++g) {
            m[g]->slab->h_status[1] = ST_RUNNING;
        }
        for (i64 b = 0;; ++b) {
            if (b >= 3) {
                bool running = true;
                for (int g = 0; g < nm; ++g) {
                    TRY(slab_wait_event(m[g], (int)((b - 3) & 3), true));
                    running = running && m[g]->slab->h_status[1] == ST_RUNNING;
                }
                if (!running) {
                    break;
                }
            }
            for (int g = 0; g < nm; ++g) {
                fqsb_system* s = m[g];
                CU(cudaSetDevice(s->device));
                RunArgs Ag = Alog;
                Ag.own_lo = (int)s->own_lo;
                Ag.own_hi = (int)s->own_hi;
                Ag.log = s->d_log;
                cudaError_t e = launch_blocked(plan[(size_t)g], s->P, s->S, Ag, s->bk, s->stream);
                if (e != cudaSuccess) {
                    return cuda_fail(e, "blocked kernel launch");
                }
                k_slab_push_blocked<<<slab_copy_grid(s->slab, 64u), 256, 0, s->stream>>>(
                    s->S, s->slab->dev, slab_planes(s), 2, s->d_log, 1);
                CU(cudaGetLastError());
                s->launches += 2;
            }
            for (int g = 0; g < nm; ++g) {
                fqsb_system* s = m[g];
                CU(cudaSetDevice(s->device));
                k_slab_import_blocked<<<slab_copy_grid(s->slab, 32u), 256, 0, s->stream>>>(
                    s->S, s->slab->dev, slab_planes(s), 2, 1, A, (int)batch);
                CU(cudaGetLastError());
                CU(cudaEventRecord(s->slab->evb[b & 3], s->stream));
                s->launches++;
                s->slab->batches++;
                s->slab->gathers++;
            }
        }
    }
    for (int g = 0; g < nm; ++g) {
        fqsb_system* s = m[g];
        CU(cudaSetDevice(s->device));
        CU(launch_blocked_settle(s->P, s->S, s->bk, s->stream));
        const unsigned rg = (unsigned)((s->R + 127) / 128);
        k_stream_settle_flags<<<rg, 128, 0, s->stream>>>(s->P, s->S);
        CU(cudaGetLastError());
        s->launches += 2;
    }
    for (int g = 0; g < nm; ++g) {
        CU(cudaSetDevice(m[g]->device));
        TRY(pull_ctl(m[g]));
        TRY(slab_status(m[g]));
        if (A.mode != MODE_FIXED) {
            m[g]->steps += m[g]->h_ctl[0].steps;
        }
    }
    return slab_flags_all_fwd(m, nm);
}

// ---- host-side StopList (GooseFEM::Iterate::StopList) in the (num, den) form
//      of the device (fqsb_device.cuh: ring_stop), any niter_tol
struct HostRing {
    std::vector<double> num, den;
    explicit HostRing(size_t n) : num(n, std::numeric_limits<double>::infinity()), den(n, 1.0) {}
    void roll_insert(double sf, double sff)
    {
        for (size_t k = 0; k + 1 < num.size(); ++k) {
            num[k] = num[k + 1];
            den[k] = den[k + 1];
        }
        num.back() = sf;
        den.back() = sff != 0.0 ? sff : 1.0; // detail.h:1516-1519
    }
    bool stop(double tol) const
    {
        const double tol2 = tol * tol, tol4 = tol2 * tol2;
        bool descending = true, less1 = true, less2 = true;
        for (size_t k = 0; k < num.size(); ++k) {
            if (k + 1 < num.size() && num[k + 1] * den[k] > num[k] * den[k + 1]) {
                descending = false;
            }
            less1 = less1 && (num[k] < tol2 * den[k]);
            less2 = less2 && (num[k] < tol4 * den[k]);
        }
        return (descending && less1) || less2;
    }
};

// replay of the per-step decisions over a batch log [k][FQSB_NLOG]; returns the 1-based step at
// which the criterion fires, 0 if it does not, -1 on NaN (detail.h:1567-1569)
static i64 slab_first_stop(const double* log, i64 k, HostRing& ring, double tol)
{
    for (i64 j = 0; j < k; ++j) {
        const double sf = log[j * FQSB_NLOG], sff = log[j * FQSB_NLOG + 1];
        if (sf != sf) {
            return -1;
        }
        ring.roll_insert(sf, sff);
        if (ring.stop(tol)) {
            return j + 1;
        }
    }
    return 0;
}

// ---- StopLists longer than the device ring (niter_tol > 32), single systems -----------------------
// timeStepsUntilEvent / minimise / minimise_truncate (detail.h:1595-1622, 1754-1792, 1833-1893) as
// batches of logged steps on the streaming kernels: snapshot, k steps whose per-step sums (and
// well-change counts) are logged on the device, replay of the reference's per-step decisions on
// the host with a StopList of any length; a stop inside a batch rolls back and redoes exactly
// that many steps. Leaves h_ctl / the device control block as the resident kernels would.
static int run_host_ring(fqsb_system* s, RunArgs A, bool overdamped, bool track_user)
{
    const i64 K = 64;
    if ((size_t)K * FQSB_NLOG > s->log_cap) {
        TRY(dev_alloc(s, &s->d_log, (size_t)K * FQSB_NLOG));
        s->log_cap = (size_t)K * FQSB_NLOG;
    }
    i64 S_run = 0, A_run = 0;
    if (track_user) { // S, A against the caller's i_n, reduced by the caller into d_out
        CU(cudaMemcpyAsync(s->h_out, s->d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost,
                           s->stream));
        CU(cudaStreamSynchronize(s->stream));
        S_run = (i64)s->h_out[2];
        A_run = (i64)s->h_out[1];
    }
    TRY(pull_ctl(s));
    const i64 inc0 = s->h_ctl[0].inc;
    i64 qs_first = overdamped ? inc0 : s->h_ctl[0].qs_first;
    i64 qs_last = overdamped ? inc0 : s->h_ctl[0].qs_last;
    int init = 1;
    i64 s_n = 0, steps = 0;
    int status = ST_RUNNING;
    double last_sf = 0.0, last_sff = 0.0;
    HostRing ring((size_t)A.niter_tol);
    std::vector<double> log((size_t)K * FQSB_NLOG);
    auto logged = [&](i64 k) -> int {
        RunArgs L = make_args(MODE_LOG, k);
        L.track = A.track;
        L.i_n = A.i_n;
        L.log = s->d_log;
        TRY(run(s, L, overdamped, false));
        CU(cudaMemcpyAsync(log.data(), s->d_log, (size_t)k * FQSB_NLOG * sizeof(double),
                           cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        return FQSB_OK;
    };
    while (status == ST_RUNNING) {
        const i64 k = A.max_steps - steps < K ? A.max_steps - steps : K;
        if (k <= 0) {
            status = ST_EXHAUSTED;
            break;
        }
        TRY(fqsb_snapshot(s));
        TRY(logged(k));
        // state of the bookkeeping at the start of the batch (a redo replays from here)
        const HostRing ring0 = ring;
        const i64 S0 = S_run, A0 = A_run, sn0 = s_n, qf0 = qs_first, ql0 = qs_last;
        const int init0 = init;
        auto replay = [&](i64 nsteps) -> i64 {
            for (i64 j = 0; j < nsteps; ++j) {
                const double* e = &log[(size_t)j * FQSB_NLOG];
                const i64 inc = inc0 + (overdamped ? 0 : steps + j + 1);
                last_sf = e[0];
                last_sff = e[1];
                if (e[0] != e[0]) { // detail.h:1567
                    status = ST_NAN;
                    return j + 1;
                }
                if (A.mode == MODE_UNTIL_EVENT && e[2] > 0.0) { // detail.h:1609
                    status = ST_EVENT;
                    return j + 1;
                }
                ring.roll_insert(e[0], e[1]);
                if (A.track) { // detail.h:1768-1778, 1863-1872
                    S_run += (i64)e[3];
                    A_run += (i64)e[4];
                    if (S_run != s_n) {
                        if (init) {
                            init = 0;
                            qs_first = inc;
                        }
                        qs_last = inc;
                    }
                    s_n = S_run;
                }
                if (ring.stop(A.tol)) { // detail.h:1615, 1780, 1874
                    status = ST_CONVERGED;
                    return j + 1;
                }
                if (A.mode == MODE_TRUNCATE) { // detail.h:1879-1885
                    if ((A.A_truncate > 0 && A_run >= A.A_truncate) ||
                        (A.S_truncate > 0 && S_run >= A.S_truncate)) {
                        status = ST_TRUNCATED;
                        return j + 1;
                    }
                }
            }
            return nsteps;
        };
        const i64 done = replay(k);
        if (status != ST_RUNNING && done < k) {
            // the call ends inside the batch: back to its start, exactly `done` steps
            TRY(fqsb_rollback(s));
            ring = ring0;
            S_run = S0;
            A_run = A0;
            s_n = sn0;
            qs_first = qf0;
            qs_last = ql0;
            init = init0;
            const int want = status;
            status = ST_RUNNING;
            TRY(logged(done));
            if (replay(done) != done || status != want) {
                return fail(FQSB_EASSERT, "host StopList: the redone batch did not reproduce its log");
            }
        }
        steps += done;
        if (status == ST_RUNNING && steps >= A.max_steps) {
            status = ST_EXHAUSTED;
        }
    }
    if (status == ST_CONVERGED) { // quench(), detail.h:1749, 1781
        k_zero_va<<<grid_for(s->n), 256, 0, s->stream>>>(s->P, s->S);
        CU(cudaGetLastError());
        s->launches++;
    }
    TRY(pull_ctl(s));
    Ctl& c = s->h_ctl[0];
    c.status = status;
    c.steps = steps;
    c.S = S_run;
    c.A = A_run;
    c.s_n = s_n;
    c.init = init;
    c.qs_first = qs_first;
    c.qs_last = qs_last;
    c.residual = last_sff != 0.0 ? std::sqrt(last_sf) / std::sqrt(last_sff) : std::sqrt(last_sf);
    TRY(push_ctl(s));
    invalidate_forces(s);
    if (status == ST_NAN) {
        return fail(FQSB_ENAN, "NaN entries found");
    }
    return check_flags(s);
}

static int slab_check_group(fqsb_system** m, int nm)
{
    if (!m || nm < 1) {
        return fail(FQSB_EASSERT, "null slab group");
    }
    for (int g = 0; g < nm; ++g) {
        TRY(slab_require(m[g]));
    }
    return FQSB_OK;
}

static int slab_flags_all(fqsb_system** m, int nm)
{
    for (int g = 0; g < nm; ++g) {
        CU(cudaSetDevice(m[g]->device));
        TRY(check_flags(m[g]));
    }
    return FQSB_OK;
}

static int slab_flags_all_fwd(fqsb_system** m, int nm) { return slab_flags_all(m, nm); }

// local reduction (k_reduce) on every member + raw gather: (*res)[world][4] on the host
static int slab_reduce_gather(fqsb_system** m, int nm, int what, int direction, bool use_mark,
                              const double** res)
{
    for (int g = 0; g < nm; ++g) {
        fqsb_system* s = m[g];
        CU(cudaSetDevice(s->device));
        if (what == 3) {
            TRY(align(s, nullptr));
        }
        if (use_mark && !s->d_mark) {
            return fail(FQSB_EASSERT, "no marked indices (fqsb_slab_mark_indices)");
        }
        TRY(reduce(s, what, direction, use_mark ? s->d_mark : nullptr));
        TRY(slab_enqueue_push(s, 4, s->d_out, 0));
    }
    for (int g = 0; g < nm; ++g) {
        CU(cudaSetDevice(m[g]->device));
        TRY(slab_enqueue_import(m[g], 4, 1, 0, g == 0 ? res : nullptr));
    }
    for (int g = 0; g < nm; ++g) {
        TRY(slab_sync(m[g]));
    }
    return FQSB_OK;
}

extern "C" {

int fqsb_slab_init(fqsb_system* s, int rank, int world, int64_t halo_cells, int kmax)
{
    TRY(enter(s));
    if (s->slab) {
        return fail(FQSB_EASSERT, "already a slab member");
    }
    if (world < 1 || world > FQSB_SLAB_MAXW || rank < 0 || rank >= world) {
        return fail(FQSB_EASSERT, ASSERT_MSG("0 <= rank < world <= 16"));
    }
    if (s->R != 1) {
        return fail(FQSB_EASSERT, ASSERT_MSG("nrealisations == 1 (a slab is part of ONE system)"));
    }
    if (halo_cells < 1 || 3 * halo_cells > s->N) {
        return fail(FQSB_EASSERT, ASSERT_MSG("1 <= halo, owned rows >= halo rows"));
    }
    // nearest-neighbour stencils with per-block disorder only: LongRange couples all blocks and
    // the thermal classes consume one ordered pcg32 stream per realisation
    if (s->P.inter == INT_LONGRANGE1D || s->thermal || s->par.minimisation == FQSB_MIN_NONE) {
        return fail(FQSB_EUNSUPPORTED,
                    "slab decomposition needs a nearest-neighbour, athermal system");
    }
    const int ksel = s->par.kernel & 15;
    if (ksel == 1) {
        return fail(FQSB_EASSERT, "a slab member cannot be forced onto the resident kernel");
    }
    // 1-D nearest-neighbour lines: the temporally blocked kernel (one launch per batch) unless the
    // streaming kernels are forced; everything else streams
    const bool blocked = ksel != 2 && s->P.rank == 1 && blocked_supported(s->P) &&
                         s->par.minimisation == FQSB_MIN_DYNAMIC && !slow_distribution(s);
    if (ksel == 3 && !blocked) {
        return fail(FQSB_EUNSUPPORTED, "no blocked kernel for this system");
    }
    if (kmax < 1) {
        kmax = 64;
    }
    fqsb_slab_state* L = new fqsb_slab_state();
    memset(L, 0, sizeof *L);
    s->slab = L;
    L->rank = rank;
    L->world = world;
    L->halo_cells = halo_cells;
    L->kmax = kmax;
    L->gcap = kmax * FQSB_NLOG > 8 ? kmax * FQSB_NLOG : 8;
    L->overdamped = s->par.minimisation == FQSB_MIN_OVERDAMPED;
    L->blocked = blocked;
    L->mailbox_bytes = slab_mailbox_words(halo_cells, world, L->gcap) * 8;
    auto build = [&]() -> int {
        CU(cudaMalloc((void**)&L->mailbox, L->mailbox_bytes));
        CU(cudaMemset(L->mailbox, 0, L->mailbox_bytes));
        CU(cudaMalloc((void**)&L->d_epoch, 2 * sizeof(u64)));
        CU(cudaMemset(L->d_epoch, 0, 2 * sizeof(u64)));
        // [0], [1]: push / import tickets; [2], [3]: readers / pushers done of a fused tile launch
        CU(cudaMalloc((void**)&L->d_ticket, 4 * sizeof(unsigned int)));
        CU(cudaMemset(L->d_ticket, 0, 4 * sizeof(unsigned int)));
        const size_t res = 2 * (size_t)world * (size_t)L->gcap;
        CU(cudaHostAlloc((void**)&L->h_res, res * sizeof(double), cudaHostAllocMapped));
        CU(cudaHostAlloc((void**)&L->h_status, 4 * sizeof(int), cudaHostAllocMapped));
        memset(L->h_status, 0, 4 * sizeof(int));
        const size_t need = (size_t)kmax * FQSB_NLOG;
        if (need > s->log_cap) {
            TRY(dev_alloc(s, &s->d_log, need));
            s->log_cap = need;
        }
        TRY(ensure_stream_buffers(s));
        for (int z = 0; z < 2; ++z) {
            if (!blocked) { // (a blocked batch is revocable by itself: no snapshots)
                for (int q = 0; q < 7; ++q) {
                    TRY(dev_alloc(s, &L->snap[z].p[q], (size_t)s->n));
                }
                TRY(dev_alloc(s, &L->snap[z].uf, 1));
                TRY(dev_alloc(s, &L->snap[z].ctl, 1));
            }
            CU(cudaEventCreateWithFlags(&L->ev[z], cudaEventDisableTiming));
        }
        for (int z = 0; z < 4; ++z) {
            CU(cudaEventCreateWithFlags(&L->evb[z], cudaEventDisableTiming));
        }
        {
            const int sweep_tiles = stream_launch_tiles(s->P, s->S.tiles, true);
            const int step_tiles = stream_launch_tiles(s->P, s->S.tiles, false);
            L->part_tiles = L->overdamped ? sweep_tiles : step_tiles;
            TRY(dev_alloc(s, &L->d_part, (size_t)kmax * (size_t)L->part_tiles * FQSB_NPART));
        }
        CU(cudaDeviceSynchronize());
        return FQSB_OK;
    };
    int rc = build();
    if (rc != FQSB_OK) {
        std::string keep = g_err;
        slab_free(s);
        g_err = keep;
        return rc;
    }
    L->peer[rank] = L->mailbox;
    s->own_lo = halo_cells;
    s->own_hi = s->N - halo_cells;
    SlabDev& D = L->dev;
    D.rank = rank;
    D.world = world;
    D.hc = halo_cells;
    D.n = s->N;
    D.gcap = L->gcap;
    D.epoch = L->d_epoch;
    D.ticket = L->d_ticket;
    D.h_res = L->h_res;
    D.h_status = L->h_status;
    D.timeout_ns = 20ULL * 1000000000ULL;
    if (const char* e = std::getenv("FQSB_SLAB_TIMEOUT_MS")) {
        D.timeout_ns = (unsigned long long)std::atoll(e) * 1000000ULL;
    }
    return FQSB_OK;
}

int fqsb_slab_ipc_handle(fqsb_system* s, void* out64)
{
    TRY(slab_require(s, false));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->slab->mailbox));
    memcpy(out64, &h, 64);
    return FQSB_OK;
}

// locals[j] != NULL: member j lives in this process (its mailbox is used directly, peer access is
// enabled between the two devices); otherwise ipc_handles + 64 * j is member j's IPC handle
int fqsb_slab_connect(fqsb_system* s, fqsb_system* const* locals, const void* ipc_handles)
{
    TRY(slab_require(s, false));
    fqsb_slab_state* L = s->slab;
    for (int j = 0; j < L->world; ++j) {
        if (j == L->rank) {
            continue;
        }
        if (locals && locals[j]) {
            fqsb_system* o = locals[j];
            if (!o->slab || o->slab->world != L->world || o->slab->rank != j ||
                o->slab->halo_cells != L->halo_cells || o->slab->gcap != L->gcap) {
                return fail(FQSB_EASSERT, "slab members do not match");
            }
            if (o->device != s->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, s->device, o->device));
                if (!can) {
                    return fail(FQSB_ECUDA, "no peer access between the devices of the slab");
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    return cuda_fail(e, "cudaDeviceEnablePeerAccess");
                }
                cudaGetLastError();
            }
            L->peer[j] = o->slab->mailbox;
            L->peer_ipc[j] = false;
        }
        else if (ipc_handles) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)ipc_handles + 64 * (size_t)j, 64);
            void* p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            L->peer[j] = (u64*)p;
            L->peer_ipc[j] = true;
        }
        else {
            return fail(FQSB_EASSERT, "slab member without a local handle or an IPC handle");
        }
    }
    for (int j = 0; j < L->world; ++j) {
        L->dev.peer[j] = L->peer[j];
    }
    L->connected = true;
    return FQSB_OK;
}

int fqsb_slab_info(fqsb_system* s, int64_t* out /* [10] */)
{
    TRY(slab_require(s, false));
    fqsb_slab_state* L = s->slab;
    out[0] = L->rank;
    out[1] = L->world;
    out[2] = L->halo_cells;
    out[3] = s->own_lo;
    out[4] = s->own_hi;
    out[5] = L->batches;
    out[6] = L->redone;
    out[8] = L->wasted;
    out[9] = L->blocked ? 1 : 0;
    out[7] = 0;
    for (int i = 0; i < FQSB_SLAB_GRAPHS; ++i) {
        out[7] += L->graphs[i].exec ? 1 : 0;
    }
    return FQSB_OK;
}

// refresh the halo rows of every member from its neighbours' owned rows
int fqsb_slab_exchange(fqsb_system** m, int nm)
{
    TRY(slab_check_group(m, nm));
    for (int g = 0; g < nm; ++g) {
        CU(cudaSetDevice(m[g]->device));
        TRY(slab_enqueue_push(m[g], 0, nullptr, 1));
    }
    for (int g = 0; g < nm; ++g) {
        CU(cudaSetDevice(m[g]->device));
        TRY(slab_enqueue_import(m[g], 0, 0, 1));
    }
    for (int g = 0; g < nm; ++g) {
        TRY(slab_sync(m[g]));
    }
    return FQSB_OK;
}

// timeSteps(n) / flowSteps(n, v_frame) of the decomposed system (detail.h:1577-1583, 1637-1645):
// batches of `batch` <= halo rows steps, one peer-memory halo exchange per batch, no host
// synchronisation until the end
int fqsb_slab_time_steps(fqsb_system** m, int nm, int64_t n, int64_t batch, int flow, double v_frame)
{
    TRY(slab_check_group(m, nm));
    for (int g = 0; g < nm; ++g) {
        TRY(require_dynamic(m[g]));
    }
    if (n < 0 || batch < 1) {
        return fail(FQSB_EASSERT, ASSERT_MSG("n >= 0 && batch >= 1"));
    }
    if (m[0]->slab->blocked) {
        RunArgs A = make_args(MODE_FIXED, n);
        A.flow = flow;
        A.v_frame = v_frame;
        return n > 0 ? slab_blocked_run(m, nm, A, batch) : FQSB_OK;
    }
    i64 left = n;
    while (left > 0) {
        const i64 k = left < batch ? left : batch;
        for (int g = 0; g < nm; ++g) {
            TRY(slab_enqueue_batch(m[g], k, MODE_FIXED, -1, flow, v_frame));
        }
        for (int g = 0; g < nm; ++g) {
            CU(cudaSetDevice(m[g]->device));
            TRY(slab_enqueue_import(m[g], 0, 0, 1));
        }
        left -= k;
    }
    for (int g = 0; g < nm; ++g) {
        TRY(slab_sync(m[g]));
    }
    return slab_flags_all(m, nm);
}

// minimise() of the decomposed system (detail.h:1676-1792, dynamic or overdamped): the reference
// decides after every step; here every batch logs its per-step global sums and the criterion is
// replayed on them. The host never sits between two batches: batch e+1 is enqueued (with its own
// snapshot) before the sums of batch e are looked at; when batch e turns out to hold the stopping
// step, the speculative batch is discarded by restoring a snapshot.
// *ret: 0 converged, steps + 1 otherwise (quirk Q4). *steps_out: steps taken.
int fqsb_slab_minimise(fqsb_system** m, int nm, double tol, int64_t niter_tol, int64_t max_iter,
                       int64_t batch, int max_iter_is_error, int64_t* ret, int64_t* steps_out)
{
    TRY(slab_check_group(m, nm));
    if (!(tol < 1.0)) {
        return fail(FQSB_EASSERT, ASSERT_MSG("tol < 1.0")); // detail.h:1684
    }
    if (niter_tol < 1 || batch < 1 || batch > m[0]->slab->kmax) {
        return fail(FQSB_EASSERT, ASSERT_MSG("niter_tol >= 1 && 1 <= batch <= kmax"));
    }
    if (m[0]->slab->blocked && niter_tol <= FQSB_RING && max_iter > 0) {
        // (longer StopLists than the device's warp-held ring take the host replay below)
        RunArgs A = make_args(MODE_MINIMISE, max_iter);
        A.tol = tol;
        A.tol2 = tol * tol;
        A.niter_tol = (int)niter_tol;
        TRY(slab_blocked_run(m, nm, A, batch));
        const Ctl& c = m[0]->h_ctl[0];
        if (c.status == ST_NAN) {
            return fail(FQSB_ENAN, "NaN entries found");
        }
        if (ret) {
            *ret = c.status == ST_CONVERGED ? 0 : c.steps + 1;
        }
        if (steps_out) {
            *steps_out = c.steps;
        }
        if (c.status != ST_CONVERGED && max_iter_is_error) {
            return fail(FQSB_ENOCONV, "No convergence found");
        }
        return FQSB_OK;
    }
    static const bool speculate = [] {
        const char* e = std::getenv("FQSB_SLAB_SPECULATE");
        return e ? std::atoi(e) != 0 : true;
    }();
    HostRing ring((size_t)niter_tol);
    // enqueue one batch on every member: snapshot into `slot`, k logged steps, push, import
    auto enqueue = [&](i64 k, int slot, const double** res) -> int {
        for (int g = 0; g < nm; ++g) {
            TRY(slab_enqueue_batch(m[g], k, MODE_LOG, slot, 0, 0.0));
        }
        for (int g = 0; g < nm; ++g) {
            CU(cudaSetDevice(m[g]->device));
            TRY(slab_enqueue_import(m[g], (int)(k * FQSB_NLOG), 0, 1, g == 0 ? res : nullptr,
                                    slot >= 0 ? slot : 0));
            m[g]->slab->batches++;
        }
        return FQSB_OK;
    };
    auto wait = [&](int slot) -> int {
        for (int g = 0; g < nm; ++g) {
            TRY(slab_wait_event(m[g], slot));
        }
        return FQSB_OK;
    };
    auto restore = [&](int slot) -> int {
        for (int g = 0; g < nm; ++g) {
            CU(cudaSetDevice(m[g]->device));
            TRY(slab_copy_state(m[g], slot, true));
        }
        return FQSB_OK;
    };
    auto finish = [&](i64 code, i64 steps) -> int {
        TRY(slab_flags_all(m, nm));
        if (ret) {
            *ret = code;
        }
        if (steps_out) {
            *steps_out = steps;
        }
        return FQSB_OK;
    };
    if (max_iter <= 0) {
        TRY(finish(1, 0));
        return max_iter_is_error ? fail(FQSB_ENOCONV, "No convergence found") : FQSB_OK;
    }
    i64 done = 0;
    int slot = 0;
    i64 k = batch < max_iter ? batch : max_iter;
    const double* res = nullptr;
    TRY(enqueue(k, slot, &res));
    for (;;) {
        // the next batch, before this one's sums are known
        const i64 left = max_iter - done - k;
        const i64 k_next = left < batch ? left : batch;
        const double* res_next = nullptr;
        const bool spec = speculate && k_next > 0;
        if (spec) {
            TRY(enqueue(k_next, slot ^ 1, &res_next));
        }
        TRY(wait(slot));
        const HostRing saved = ring;
        const i64 stop = slab_first_stop(res, k, ring, tol);
        if (stop < 0) {
            for (int g = 0; g < nm; ++g) {
                slab_sync(m[g]);
            }
            return fail(FQSB_ENAN, "NaN entries found"); // detail.h:1568
        }
        if (stop > 0) {
            if (stop < k) { // the criterion fired inside the batch: redo exactly `stop` steps
                TRY(restore(slot));
                ring = saved;
                const double* res_redo = nullptr;
                TRY(enqueue(stop, -1, &res_redo));
                TRY(wait(0));
                for (int g = 0; g < nm; ++g) {
                    m[g]->slab->redone++;
                }
                if (slab_first_stop(res_redo, stop, ring, tol) != stop) {
                    return fail(FQSB_EASSERT, "slab: the redone batch did not reproduce its log");
                }
            }
            else if (spec) { // the speculative batch started from exactly the state to keep
                TRY(restore(slot ^ 1));
            }
            if (spec) {
                for (int g = 0; g < nm; ++g) {
                    m[g]->slab->wasted++;
                }
            }
            for (int g = 0; g < nm; ++g) { // quench(), detail.h:1781
                CU(cudaSetDevice(m[g]->device));
                k_zero_va<<<grid_for(m[g]->n), 256, 0, m[g]->stream>>>(m[g]->P, m[g]->S);
                CU(cudaGetLastError());
                m[g]->launches++;
            }
            return finish(0, done + stop);
        }
        done += k;
        if (k_next <= 0) {
            break;
        }
        if (!spec) {
            TRY(enqueue(k_next, slot ^ 1, &res_next));
        }
        k = k_next;
        slot ^= 1;
        res = res_next;
    }
    TRY(finish(done + 1, done)); // detail.h:1791 (quirk Q4)
    if (max_iter_is_error) {
        return fail(FQSB_ENOCONV, "No convergence found"); // detail.h:1788
    }
    return FQSB_OK;
}

// sums over the whole decomposed system, members added in rank order. out[4]:
//   what 1: {sum f^2, sum f_frame^2}   2: {sum v^2, sum f_frame}
//   what 3: {-, off-branch count, -, min displacement}
//   what 4: {sum (i - i_mark), #(i != i_mark), sum |i - i_mark|} against fqsb_slab_mark_indices
int fqsb_slab_sums(fqsb_system** m, int nm, int what, int direction, double* out)
{
    TRY(slab_check_group(m, nm));
    if (what < 1 || what > 4) {
        return fail(FQSB_EASSERT, "unknown reduction");
    }
    const double* res = nullptr;
    TRY(slab_reduce_gather(m, nm, what, direction, what == 4, &res));
    const fqsb_slab_state* L = m[0]->slab;
    double acc[3] = {0.0, 0.0, 0.0}, mn = 1.7976931348623157e308;
    for (int j = 0; j < L->world; ++j) {
        for (int c = 0; c < 3; ++c) {
            acc[c] += res[4 * j + c];
        }
        mn = std::fmin(mn, res[4 * j + 3]);
    }
    out[0] = acc[0];
    out[1] = acc[1];
    out[2] = acc[2];
    out[3] = mn;
    return what == 3 ? slab_flags_all(m, nm) : FQSB_OK;
}

// device-side copy of the current well indices: the i_n of the examples' S = sum(i - i_n)
int fqsb_slab_mark_indices(fqsb_system** m, int nm)
{
    TRY(slab_check_group(m, nm));
    for (int g = 0; g < nm; ++g) {
        CU(cudaSetDevice(m[g]->device));
        TRY(mark_index(m[g]));
    }
    return FQSB_OK;
}

// eventDrivenStep(eps, kick, direction) of the decomposed system (detail.h:1933-1960): the uniform
// displacement is agreed over all members (one gather of 4 doubles), applied locally -- halo rows
// move with their originals, no exchange needed
int fqsb_slab_event_driven_step(fqsb_system** m, int nm, double eps, int kick, int direction,
                                double* du_frame)
{
    TRY(slab_check_group(m, nm));
    if (direction != 1 && direction != -1) {
        return fail(FQSB_EASSERT, ASSERT_MSG("direction == 1 || direction == -1"));
    }
    const Par& P = m[0]->P;
    double dup;
    if (!kick) {
        if (P.pot == POT_SMOOTH) {
            return fail(FQSB_EUNSUPPORTED, "Operation not possible."); // detail.h:420
        }
        double sums[4];
        TRY(fqsb_slab_sums(m, nm, 3, direction, sums));
        const double d = sums[1] > 0.0 ? 0.0 : sums[3];
        if (d < 0.5 * eps) {
            if (du_frame) {
                *du_frame = 0.0;
            }
            return FQSB_OK;
        }
        dup = direction > 0 ? d - 0.5 * eps : 0.5 * eps - d;
    }
    else {
        dup = direction > 0 ? eps : -eps;
    }
    const double duf = dup * (P.k_frame + P.mu) / P.k_frame;
    for (int g = 0; g < nm; ++g) {
        fqsb_system* s = m[g];
        CU(cudaSetDevice(s->device));
        k_slab_shift<<<1, 1, 0, s->stream>>>(s->S, s->d_du, dup, duf);
        TRY(advance(s));
    }
    for (int g = 0; g < nm; ++g) {
        CU(cudaSetDevice(m[g]->device));
        TRY(check_flags(m[g]));
    }
    if (du_frame) {
        *du_frame = duf;
    }
    return FQSB_OK;
}

// the StopList replay as a plain host function (unit-tested without a GPU): ring_num / ring_den
// [niter_tol] carry the state between calls (start: +inf / 1)
int64_t fqsb_slab_first_stop(const double* log, int64_t k, double tol, int64_t niter_tol,
                             double* ring_num, double* ring_den)
{
    HostRing ring((size_t)niter_tol);
    for (int64_t i = 0; i < niter_tol; ++i) {
        ring.num[(size_t)i] = ring_num[i];
        ring.den[(size_t)i] = ring_den[i];
    }
    const i64 stop = slab_first_stop(log, k, ring, tol);
    for (int64_t i = 0; i < niter_tol; ++i) {
        ring_num[i] = ring.num[(size_t)i];
        ring_den[i] = ring.den[(size_t)i];
    }
    return stop;
}

} // extern "C"
