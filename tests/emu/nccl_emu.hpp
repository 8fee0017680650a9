// TEST SUPPORT ONLY — included by slab.cu when it is compiled by a plain C++ compiler for the lock-step host replay (never by
// nvcc, never by the product build).  An in-process stand-in for the handful of NCCL calls the slab decomposition makes:
// the ranks of a "node" are OS threads of the test process, a communicator is a shared mailbox set.
//   - point-to-point: ncclSend / ncclRecv between ncclGroupStart / ncclGroupEnd; messages between one (source, destination) pair
//     are matched in posting order, as NCCL does (with two ranks the left and the right neighbour are the same peer and the
//     order is what tells the two messages apart);
//   - ncclAllReduce: sum / max / min over the ranks in rank order (deterministic), every rank gets the result;
//   - all calls complete before they return (the replay's "streams" are synchronous).
#pragma once
#include <string.h>

#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#define NCCL_UNIQUE_ID_BYTES 128
typedef struct { char internal[NCCL_UNIQUE_ID_BYTES]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclChar = 0, ncclInt = 2, ncclUint64 = 5, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;

struct EmuNcclWorld {
    int n = 0, joined = 0;
    std::mutex m;
    std::condition_variable cv;
    std::vector<std::deque<std::vector<char>>> box;  // [src * n + dst]
    int arrived = 0;
    unsigned long long generation = 0;
    std::vector<std::vector<char>> contrib;
    std::vector<char> result;
};
struct EmuNcclComm { EmuNcclWorld* w; int rank; };
typedef EmuNcclComm* ncclComm_t;

struct EmuNcclOp { bool send; const void* sbuf; void* rbuf; size_t bytes; int peer; EmuNcclComm* comm; };
inline std::mutex emu_nccl_registry_m;
inline std::map<std::string, EmuNcclWorld*> emu_nccl_registry;
inline unsigned long long emu_nccl_next_id = 1;
inline thread_local int emu_nccl_group_depth = 0;
inline thread_local std::vector<EmuNcclOp> emu_nccl_pending;

static inline size_t emu_nccl_size(ncclDataType_t t) { return t == ncclChar ? 1 : t == ncclInt ? 4 : 8; }

static inline ncclResult_t emu_ncclGetUniqueId(ncclUniqueId* id) {
    std::lock_guard<std::mutex> lock(emu_nccl_registry_m);
    memset(id->internal, 0, NCCL_UNIQUE_ID_BYTES);
    snprintf(id->internal, NCCL_UNIQUE_ID_BYTES, "pfmds-host-replay-%llu", emu_nccl_next_id++);
    return ncclSuccess;
}
static inline ncclResult_t emu_ncclCommInitRank(ncclComm_t* comm, int n, ncclUniqueId id, int rank) {
    EmuNcclWorld* w;
    {
        std::lock_guard<std::mutex> lock(emu_nccl_registry_m);
        std::string key(id.internal, strnlen(id.internal, NCCL_UNIQUE_ID_BYTES));
        EmuNcclWorld*& slot = emu_nccl_registry[key];
        if (!slot) { slot = new EmuNcclWorld; slot->n = n; slot->box.resize((size_t)n * n); slot->contrib.resize((size_t)n); }
        w = slot;
    }
    if (w->n != n || rank < 0 || rank >= n) return ncclInvalidArgument;
    *comm = new EmuNcclComm{w, rank};
    std::unique_lock<std::mutex> lock(w->m);   // every rank has joined before anyone communicates
    w->joined += 1;
    w->cv.notify_all();
    w->cv.wait(lock, [&] { return w->joined >= w->n; });
    return ncclSuccess;
}
static inline ncclResult_t emu_ncclCommDestroy(ncclComm_t comm) { delete comm; return ncclSuccess; }   // the world is left to the process
static inline void emu_nccl_run(const EmuNcclOp& op) {
    EmuNcclWorld* w = op.comm->w;
    const int me = op.comm->rank;
    std::unique_lock<std::mutex> lock(w->m);
    if (op.send) {
        const char* b = (const char*)op.sbuf;
        w->box[(size_t)me * w->n + op.peer].emplace_back(b, b + op.bytes);
        w->cv.notify_all();
    } else {
        auto& q = w->box[(size_t)op.peer * w->n + me];
        w->cv.wait(lock, [&] { return !q.empty(); });
        std::vector<char> msg = std::move(q.front());
        q.pop_front();
        if (msg.size() != op.bytes) { fprintf(stderr, "host replay NCCL: rank %d expected %zu bytes from %d, got %zu\n", me, op.bytes, op.peer, msg.size()); abort(); }
        memcpy(op.rbuf, msg.data(), op.bytes);
    }
}
static inline ncclResult_t emu_ncclGroupStart() { emu_nccl_group_depth += 1; return ncclSuccess; }
static inline ncclResult_t emu_ncclGroupEnd() {
    if (--emu_nccl_group_depth > 0) return ncclSuccess;
    std::vector<EmuNcclOp> ops;
    ops.swap(emu_nccl_pending);
    for (auto& op : ops) if (op.send) emu_nccl_run(op);    // all sends first: nothing in a group may wait for a peer's receive
    for (auto& op : ops) if (!op.send) emu_nccl_run(op);
    return ncclSuccess;
}
static inline ncclResult_t emu_ncclSend(const void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t) {
    EmuNcclOp op{true, buf, nullptr, count * emu_nccl_size(t), peer, comm};
    if (emu_nccl_group_depth > 0) emu_nccl_pending.push_back(op); else emu_nccl_run(op);
    return ncclSuccess;
}
static inline ncclResult_t emu_ncclRecv(void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t) {
    EmuNcclOp op{false, nullptr, buf, count * emu_nccl_size(t), peer, comm};
    if (emu_nccl_group_depth > 0) emu_nccl_pending.push_back(op); else emu_nccl_run(op);
    return ncclSuccess;
}
template <class T>
static inline void emu_nccl_reduce(EmuNcclWorld* w, size_t count, ncclRedOp_t op) {
    w->result.assign(count * sizeof(T), 0);
    T* out = (T*)w->result.data();
    for (int r = 0; r < w->n; ++r) {
        const T* in = (const T*)w->contrib[(size_t)r].data();
        for (size_t k = 0; k < count; ++k) {
            if (r == 0) out[k] = in[k];
            else if (op == ncclSum) out[k] = out[k] + in[k];
            else if (op == ncclMax) out[k] = in[k] > out[k] ? in[k] : out[k];
            else out[k] = in[k] < out[k] ? in[k] : out[k];
        }
    }
}
static inline ncclResult_t emu_ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t comm, cudaStream_t) {
    EmuNcclWorld* w = comm->w;
    const size_t bytes = count * emu_nccl_size(t);
    std::unique_lock<std::mutex> lock(w->m);
    const unsigned long long gen = w->generation;
    const char* b = (const char*)send;
    w->contrib[(size_t)comm->rank].assign(b, b + bytes);
    if (++w->arrived == w->n) {
        if (t == ncclDouble) emu_nccl_reduce<double>(w, count, op);
        else if (t == ncclInt) emu_nccl_reduce<int>(w, count, op);
        else if (t == ncclUint64) emu_nccl_reduce<unsigned long long>(w, count, op);
        else emu_nccl_reduce<char>(w, count, op);
        w->arrived = 0;
        w->generation += 1;
        w->cv.notify_all();
    } else {
        w->cv.wait(lock, [&] { return w->generation != gen; });
    }
    memcpy(recv, w->result.data(), bytes);   // the next reduction cannot complete before this rank has entered it
    return ncclSuccess;
}
static inline const char* emu_ncclGetErrorString(ncclResult_t) { return "host replay NCCL stand-in"; }
