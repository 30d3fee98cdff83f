// batcher.cu -- the micro-batcher of the drop-in boundary: the reference calls
// Scheduler.Schedule(ctx, *InferenceRequest, []Endpoint) once per in-flight HTTP request, concurrently from one
// goroutine each (pkg/epp/requestcontrol/director.go:243, handlers/server.go:168); the engine decides batches.  This
// layer turns the one into the other INSIDE the library, so that a cgo shim is a two-call wrapper
// (epp_submit + epp_wait) with no batching logic of its own:
//
//   epp_submit   reserves a row of the pinned staging buffer that is currently being filled (under the lock), copies the
//                prompt into it (outside the lock) and returns a ticket;
//   the flusher  (one thread per batcher) closes a batch when it holds max_batch requests or its oldest request has
//                waited max_delay_us, runs epp_schedule on it (ONE frozen snapshot per flush, App. A.8), then -- when
//                index_picks is set -- epp_index_add_picked (PreRequest, approximateprefix/plugin.go:164-200), so index
//                updates land BETWEEN flushes; meanwhile the submitters fill the next staging buffer;
//   epp_wait     returns the ticket's decision once its batch has been flushed: it spins on the flush counter for a few
//                tens of microseconds (a flush is ONE 30-130 us kernel launch) before it blocks on the condition variable.
//
// It is built on the public C ABI only (include/epp_engine.h): no engine internals, no CPU compute path.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/epp_engine.h"

namespace {
constexpr int kBufs = 2;                       // staging buffers: one being filled while the other is in flight
constexpr int kResults = 4096;                 // flushed batches whose results are kept for epp_wait (ring)
using Clock = std::chrono::steady_clock;

struct Staging {
    uint8_t *data = nullptr;                   // [max_batch][row_cap] pinned
    uint64_t *offsets = nullptr;               // [max_batch + 1]
    uint64_t *lengths = nullptr;               // [max_batch] true prompt lengths (may exceed the row)
    uint32_t *model_ids = nullptr;
    uint8_t *multimodal = nullptr;
    epp_decision *dec = nullptr;
    epp_decision_detail *det = nullptr;
    uint32_t *topk[3] = {nullptr, nullptr, nullptr};   // [max_batch][pick_k] primary / prefill / encode lists (pick_k > 1)
    int32_t n = 0;
};
// Results of one flushed batch, kept until every ticket has been waited for or kResults later batches were flushed.
struct Result {
    uint64_t seq = ~0ull;
    int32_t status = EPP_OK;                   // return code of the batch's epp_schedule / epp_index_add_picked
    int32_t remaining = 0;                     // tickets not yet waited for
    std::vector<epp_decision> dec;
    std::vector<epp_decision_detail> det;
    std::vector<uint32_t> topk;                // [3][n][pick_k] when the engine's pick_k > 1
};
}  // namespace

struct epp_batcher {
    epp_engine *eng = nullptr;
    epp_batcher_cfg cfg{};
    uint64_t row_cap = 0;                      // bytes of a prompt that hashPrompt can read (hashing.go:63-66)
    int32_t pick_k = 0;                        // the engine's maxNumOfEndpoints (> 1: flushes also fetch the first-k lists)
    Staging buf[kBufs];
    std::vector<Result> results;
    bool in_flight[kBufs] = {false, false};    // the flusher is running epp_schedule on the buffer
    std::atomic<int32_t> writers[kBufs];       // submitters still copying their prompt into the buffer (outside the lock)
    std::atomic<uint64_t> done_seq{0};         // = seq_done, readable without the lock (epp_wait spins on it)
    std::atomic<int32_t> sleepers{0};          // waiters blocked on cv_done
    std::atomic<int32_t> spinners{0};          // waiters polling done_seq
    std::mutex mu;
    std::condition_variable cv_flush, cv_done, cv_space;
    int cur = 0;                               // buffer being filled
    int32_t n_cur = 0;
    uint64_t seq_filling = 0, seq_done = 0;    // batches [0, seq_done) have been flushed
    Clock::time_point first_arrival;
    bool stop = false;
    std::thread flusher;
    uint64_t n_flushes = 0, n_requests = 0, n_full = 0;
    std::string last_error;
};

static thread_local std::string g_batcher_error;

static void flusher_main(epp_batcher *b) {
    std::unique_lock<std::mutex> lk(b->mu);
    for (;;) {
        // wait for a batch to close: full, or its oldest request has waited max_delay_us
        for (;;) {
            if (b->stop && b->n_cur == 0) return;
            if (b->n_cur >= b->cfg.max_batch || (b->n_cur > 0 && b->stop)) break;
            if (b->n_cur > 0) {
                const auto deadline = b->first_arrival + std::chrono::microseconds(b->cfg.max_delay_us);
                if (Clock::now() >= deadline) break;
                b->cv_flush.wait_until(lk, deadline);
            } else {
                b->cv_flush.wait(lk);
            }
        }
        Staging &st = b->buf[b->cur];
        const int32_t n = b->n_cur;
        const uint64_t seq = b->seq_filling;
        st.n = n;
        if (n >= b->cfg.max_batch) b->n_full++;
        const int mine = b->cur;
        b->in_flight[mine] = true;
        b->cur = (b->cur + 1) % kBufs;          // the other buffer is free: its flush completed before this one began
        b->n_cur = 0;
        b->seq_filling++;
        lk.unlock();
        b->cv_space.notify_all();
        while (b->writers[mine].load(std::memory_order_acquire) != 0) {}      // rows reserved before the swap: a 16 KiB memcpy each

        epp_batch batch;
        memset(&batch, 0, sizeof batch);
        batch.n_requests = n;
        batch.data = st.data;
        batch.offsets = st.offsets;
        batch.lengths = st.lengths;
        batch.model_ids = st.model_ids;
        batch.multimodal = st.multimodal;
        batch.flags = EPP_BATCH_LENGTHS_EXCEED_ROWS;
        int32_t rc;
        if (b->pick_k > 1) {
            epp_topk_out tk;
            memset(&tk, 0, sizeof tk);
            tk.struct_size = sizeof tk;
            tk.k = b->pick_k;
            tk.primary = st.topk[0]; tk.prefill = st.topk[1]; tk.encode = st.topk[2];
            rc = epp_schedule_topk(b->eng, &batch, st.dec, st.det, b->cfg.index_picks ? 1 : 0, &tk);
        } else {
            rc = epp_schedule(b->eng, &batch, st.dec, st.det, b->cfg.index_picks ? 1 : 0);
        }
        std::string err;
        if (rc != EPP_OK) err = epp_last_error();
        if (rc == EPP_OK && b->cfg.index_picks) {
            const int32_t rc2 = epp_index_add_picked(b->eng);      // PreRequest: visible to the NEXT flush
            if (rc2 != EPP_OK) { rc = rc2; err = epp_last_error(); }
        }
        lk.lock();
        Result &res = b->results[seq % kResults];
        res.seq = seq;
        res.status = rc;
        res.remaining = n;
        res.dec.assign(st.dec, st.dec + n);                   // the vectors keep their capacity from round to round of the ring
        res.det.assign(st.det, st.det + n);
        res.topk.clear();
        if (b->pick_k > 1)
            for (int q = 0; q < 3; q++) res.topk.insert(res.topk.end(), st.topk[q], st.topk[q] + (size_t)n * (size_t)b->pick_k);
        b->in_flight[mine] = false;
        if (rc != EPP_OK) b->last_error = err;
        b->seq_done = seq + 1;
        b->done_seq.store(seq + 1, std::memory_order_release);
        b->n_flushes++;
        b->n_requests += (uint64_t)n;
        if (b->sleepers.load(std::memory_order_acquire) > 0) b->cv_done.notify_all();
        b->cv_space.notify_all();
    }
}

static int32_t bfail(int32_t code, const char *msg) {
    g_batcher_error = msg;
    return code;
}

extern "C" const char *epp_batcher_last_error(void) { return g_batcher_error.c_str(); }

extern "C" int32_t epp_batcher_create(epp_engine *h, const epp_batcher_cfg *cfg, epp_batcher **out) {
    if (!h || !cfg || !out) return bfail(EPP_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(epp_batcher_cfg)) return bfail(EPP_ERR_INVALID, "epp_batcher_cfg.struct_size mismatch");
    if (cfg->max_batch <= 0 || cfg->max_batch > (1 << 20)) return bfail(EPP_ERR_INVALID, "max_batch out of range [1, 2^20]");
    if (cfg->max_delay_us < 0) return bfail(EPP_ERR_INVALID, "max_delay_us must be >= 0");
    epp_config ec;
    int32_t rc = epp_get_config(h, &ec);
    if (rc != EPP_OK) return bfail(rc, epp_last_error());
    epp_batcher *b = new epp_batcher();
    b->eng = h;
    b->cfg = *cfg;
    // every byte hashPrompt can read: maxPrefixBlocks blocks of blockSizeTokens * 4 bytes (hashing.go:49, 63-66);
    // rows start on 32-byte boundaries (selects the aligned hash kernels)
    b->row_cap = ((uint64_t)ec.max_prefix_blocks * (uint64_t)ec.block_size_tokens * 4 + 31) & ~31ull;
    b->pick_k = ec.pick_k;
    b->results.resize(kResults);
    for (int i = 0; i < kBufs; i++) b->writers[i].store(0);
    const size_t mb = (size_t)cfg->max_batch;
    for (int i = 0; i < kBufs; i++) {
        Staging &st = b->buf[i];
        void *p = nullptr;
        const size_t sizes[7] = {mb * b->row_cap, sizeof(uint64_t) * (mb + 1), sizeof(uint64_t) * mb, sizeof(uint32_t) * mb,
                                 mb, sizeof(epp_decision) * mb, sizeof(epp_decision_detail) * mb};
        void **dst[7] = {(void **)&st.data, (void **)&st.offsets, (void **)&st.lengths, (void **)&st.model_ids,
                         (void **)&st.multimodal, (void **)&st.dec, (void **)&st.det};
        for (int k = 0; k < 7; k++) {
            rc = epp_host_alloc(sizes[k], &p);
            if (rc != EPP_OK) {
                g_batcher_error = epp_last_error();
                epp_batcher_destroy(b);
                return rc;
            }
            *dst[k] = p;
        }
        for (size_t r = 0; r <= mb; r++) st.offsets[r] = (uint64_t)r * b->row_cap;
        if (b->pick_k > 1)
            for (int q = 0; q < 3; q++) {
                rc = epp_host_alloc(sizeof(uint32_t) * mb * (size_t)b->pick_k, &p);
                if (rc != EPP_OK) {
                    g_batcher_error = epp_last_error();
                    epp_batcher_destroy(b);
                    return rc;
                }
                st.topk[q] = (uint32_t *)p;
            }
    }
    b->flusher = std::thread(flusher_main, b);
    *out = b;
    return EPP_OK;
}

extern "C" int32_t epp_batcher_destroy(epp_batcher *b) {
    if (!b) return EPP_OK;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->stop = true;
    }
    b->cv_flush.notify_all();
    b->cv_space.notify_all();
    if (b->flusher.joinable()) b->flusher.join();            // drains the batch being filled
    b->cv_done.notify_all();
    for (int i = 0; i < kBufs; i++) {
        Staging &st = b->buf[i];
        void *ps[10] = {st.data, st.offsets, st.lengths, st.model_ids, st.multimodal, st.dec, st.det, st.topk[0], st.topk[1], st.topk[2]};
        for (void *p : ps) if (p) epp_host_free(p);
    }
    delete b;
    return EPP_OK;
}

extern "C" int32_t epp_submit(epp_batcher *b, uint32_t model_id, const void *prompt, uint64_t prompt_len,
                              uint32_t multimodal, uint64_t *out_ticket) {
    if (!b || !out_ticket || (prompt_len && !prompt)) return bfail(EPP_ERR_INVALID, "NULL argument");
    std::unique_lock<std::mutex> lk(b->mu);
    // full: the flusher is about to take this buffer; in flight: the previous flush still reads the buffer we would fill
    while (!b->stop && (b->n_cur >= b->cfg.max_batch || b->in_flight[b->cur])) b->cv_space.wait(lk);
    if (b->stop) return bfail(EPP_ERR_STATE, "the batcher is shutting down");
    const int mine = b->cur;
    Staging &st = b->buf[mine];
    const int32_t idx = b->n_cur;
    st.lengths[idx] = prompt_len;                  // the TRUE length: the P/D decider counts it (prefix_based_pd_decider.go:152-167)
    st.model_ids[idx] = model_id;
    st.multimodal[idx] = multimodal ? 1 : 0;
    if (idx == 0) b->first_arrival = Clock::now();
    b->n_cur = idx + 1;
    *out_ticket = (b->seq_filling << 20) | (uint64_t)idx;
    const bool wake = idx == 0 || b->n_cur >= b->cfg.max_batch;
    b->writers[mine].fetch_add(1, std::memory_order_acq_rel);      // the flusher waits for this copy before it reads the row
    lk.unlock();
    if (wake) b->cv_flush.notify_one();
    const uint64_t n_copy = prompt_len < b->row_cap ? prompt_len : b->row_cap;
    if (n_copy) memcpy(st.data + (size_t)idx * b->row_cap, prompt, n_copy);
    b->writers[mine].fetch_sub(1, std::memory_order_release);
    return EPP_OK;
}

extern "C" int32_t epp_wait(epp_batcher *b, uint64_t ticket, epp_decision *out, epp_decision_detail *out_detail) {
    return epp_wait_topk(b, ticket, out, out_detail, nullptr, nullptr, nullptr);
}

extern "C" int32_t epp_wait_topk(epp_batcher *b, uint64_t ticket, epp_decision *out, epp_decision_detail *out_detail,
                                 uint32_t *primary, uint32_t *prefill, uint32_t *encode) {
    if (!b || !out) return bfail(EPP_ERR_INVALID, "NULL argument");
    if ((primary || prefill || encode) && b->pick_k <= 1) return bfail(EPP_ERR_STATE, "the engine was created with pick_k <= 1: there are no lists");
    const uint64_t seq = ticket >> 20;
    const uint32_t idx = (uint32_t)(ticket & 0xFFFFFu);
    // a flush is one 30-130 us launch: up to four waiters spin on the flush counter for about that long before they
    // block (more spinners than that only take cores away from the flusher and from each other)
    if (b->spinners.fetch_add(1, std::memory_order_acq_rel) < 4) {
        for (int spin = 0; spin < 4000 && b->done_seq.load(std::memory_order_acquire) <= seq; spin++) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    b->spinners.fetch_sub(1, std::memory_order_acq_rel);
    std::unique_lock<std::mutex> lk(b->mu);
    if (seq > b->seq_filling || (seq == b->seq_filling && (int32_t)idx >= b->n_cur)) return bfail(EPP_ERR_INVALID, "unknown ticket");
    while (b->seq_done <= seq) {                             // a closing batcher still flushes what is pending
        b->sleepers.fetch_add(1, std::memory_order_acq_rel);
        b->cv_done.wait(lk);
        b->sleepers.fetch_sub(1, std::memory_order_acq_rel);
    }
    Result &res = b->results[seq % kResults];
    if (res.seq != seq) return bfail(EPP_ERR_STATE, "ticket expired: its batch was flushed more than 4096 flushes ago, or every ticket of it was already waited for");
    if ((size_t)idx >= res.dec.size()) return bfail(EPP_ERR_INVALID, "unknown ticket");
    int32_t rc = res.status;
    if (rc != EPP_OK) {
        g_batcher_error = b->last_error;
    } else {
        *out = res.dec[idx];
        if (out_detail) *out_detail = res.det[idx];
        uint32_t *dst[3] = {primary, prefill, encode};
        const size_t k = (size_t)b->pick_k, n = res.dec.size();
        for (int q = 0; q < 3; q++)
            if (dst[q]) memcpy(dst[q], res.topk.data() + ((size_t)q * n + idx) * k, sizeof(uint32_t) * k);
    }
    if (--res.remaining == 0) {                    // every ticket served: release the memory early
        res.seq = ~0ull;
        if (res.dec.capacity() > 4096) {            // big batches give their memory back, small ones keep it for the next lap
            std::vector<epp_decision>().swap(res.dec);
            std::vector<epp_decision_detail>().swap(res.det);
            std::vector<uint32_t>().swap(res.topk);
        }
    }
    return rc;
}

extern "C" int32_t epp_batcher_stats(epp_batcher *b, epp_batcher_stats_t *out) {
    if (!b || !out) return bfail(EPP_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(b->mu);
    out->n_flushes = b->n_flushes;
    out->n_requests = b->n_requests;
    out->n_full_flushes = b->n_full;
    out->n_pending = (uint64_t)b->n_cur;
    return EPP_OK;
}
