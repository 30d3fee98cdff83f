// Package eppcuda is the cgo shim between the llm-d EPP (pkg/epp) and libepp_engine.so.
//
// It is shipped as SOURCE ONLY: the build environment of this repository has no Go toolchain, so this file has not
// been compiled.  It binds exactly the symbols declared in include/epp_engine.h and keeps the reference's plugin
// interfaces unchanged:
//
//	Scheduler.Schedule        pkg/epp/scheduling/scheduler.go:54        -> epp_schedule (micro-batched)
//	DataProducer.Produce      framework/interface/requestcontrol/plugins.go:69-72 -> epp_prefix_match
//	Scorer.Score              framework/interface/scheduling/plugins.go:68-72      -> epp_score
//	PreRequest.PreRequest     framework/interface/requestcontrol/plugins.go:36-39  -> epp_index_add
//
// Build (once Go is available):  CGO_ENABLED=1 CGO_CFLAGS=-I${REPO}/include CGO_LDFLAGS="-L${REPO}/llm-d-inference-scheduler_b200 -lepp_engine" go build ./...
package eppcuda

/*
#cgo LDFLAGS: -lepp_engine
#include <stdlib.h>
#include "epp_engine.h"
*/
import "C"

import (
	"context"
	"errors"
	"fmt"
	"sync"
	"time"
	"unsafe"
)

// Engine wraps one epp_engine handle (one GPU).
type Engine struct {
	h     *C.epp_engine
	mu    sync.Mutex        // guards slots
	slots map[string]uint32 // NamespacedName -> dense slot id
	free  []uint32
	cfg   C.epp_config
	batch *batcher
}

func lastErr(code C.int32_t) error {
	return fmt.Errorf("epp_engine error %d: %s", int(code), C.GoString(C.epp_last_error()))
}

// ScorerSpec mirrors one `pluginRef` + weight of a SchedulingProfile (configloader.go:229-247).
type ScorerSpec struct {
	Kind   int     // C.EPP_SCORER_*
	Weight float64 // default 1.0 (loader/defaults.go:42)
	Param  float64
	Column int     // ext column read by EPP_SCORER_TOKEN_LOAD / EPP_SCORER_ACTIVE_REQUEST
	Param2 float64 // active-request: idleThreshold
}

type ProfileSpec struct {
	Filter  int // C.EPP_FILTER_*
	Scorers []ScorerSpec
}

type Config struct {
	Device               int
	MaxEndpoints         int
	BlockSizeTokens      int // approximateprefix config.BlockSizeTokens
	MaxPrefixBlocks      int // config.MaxPrefixBlocksToMatch
	LRUCapacityPerServer int
	Primary              ProfileSpec  // the single profile, or "decode"
	Prefill              *ProfileSpec // non-nil => disagg-profile-handler
	NonCachedTokens      int64        // prefix-based-pd-decider
	MaxBatch             int          // micro-batcher: flush at this many requests ...
	MaxDelay             time.Duration // ... or after this long (e.g. 200us)
}

func fillProfile(dst *C.epp_profile_cfg, p ProfileSpec) {
	dst.filter = C.int32_t(p.Filter)
	dst.n_scorers = C.int32_t(len(p.Scorers))
	for i, s := range p.Scorers {
		dst.scorers[i].kind = C.int32_t(s.Kind)
		dst.scorers[i].weight = C.double(s.Weight)
		dst.scorers[i].param = C.double(s.Param)
		dst.scorers[i].column = C.int32_t(s.Column)
		dst.scorers[i].param2 = C.double(s.Param2)
	}
}

// New creates the engine (plugin factory; registry.go:25-30).
func New(c Config) (*Engine, error) {
	e := &Engine{slots: map[string]uint32{}}
	C.epp_config_default(&e.cfg)
	e.cfg.device = C.int32_t(c.Device)
	e.cfg.max_endpoints = C.int32_t(c.MaxEndpoints)
	e.cfg.block_size_tokens = C.int32_t(c.BlockSizeTokens)
	e.cfg.max_prefix_blocks = C.int32_t(c.MaxPrefixBlocks)
	e.cfg.lru_capacity_per_server = C.int32_t(c.LRUCapacityPerServer)
	fillProfile(&e.cfg.primary, c.Primary)
	if c.Prefill != nil {
		e.cfg.handler = C.EPP_HANDLER_DISAGG
		e.cfg.non_cached_tokens = C.int64_t(c.NonCachedTokens)
		fillProfile(&e.cfg.prefill, *c.Prefill)
	}
	if rc := C.epp_engine_create(&e.cfg, &e.h); rc != 0 {
		return nil, lastErr(rc) // EPP_ERR_NO_DEVICE: keep the stock Go scorers registered instead
	}
	e.batch = newBatcher(e, c.MaxBatch, c.MaxDelay)
	return e, nil
}

func (e *Engine) Close() { C.epp_engine_destroy(e.h) }

// RegisterModel registers request.TargetModel (+ cache salt); returns the id used in batches.
func (e *Engine) RegisterModel(model, salt string) (uint32, error) {
	var id C.uint32_t
	m, s := []byte(model), []byte(salt)
	var mp, sp *C.uint8_t
	if len(m) > 0 {
		mp = (*C.uint8_t)(unsafe.Pointer(&m[0]))
	}
	if len(s) > 0 {
		sp = (*C.uint8_t)(unsafe.Pointer(&s[0]))
	}
	if rc := C.epp_model_register(e.h, mp, C.size_t(len(m)), sp, C.size_t(len(s)), &id); rc != 0 {
		return 0, lastErr(rc)
	}
	return uint32(id), nil
}

// PoolEntry is what the 50 ms scrape loop knows about one endpoint (fwkdl.Metrics + role label).
type PoolEntry struct {
	Name        string // NamespacedName
	Role        uint8  // C.EPP_ROLE_*
	KVUsage     float64
	Waiting     int32
	Running     int32
}

// SetPool replaces the snapshot (Datastore.PodList + metrics; director.go:222-230).
func (e *Engine) SetPool(entries []PoolEntry) error {
	n := len(entries)
	ids := make([]C.uint32_t, n)
	role := make([]C.uint8_t, n)
	kv := make([]C.double, n)
	wq := make([]C.int32_t, n)
	rq := make([]C.int32_t, n)
	e.mu.Lock()
	for i, p := range entries {
		ids[i] = C.uint32_t(e.slotOf(p.Name))
		role[i], kv[i], wq[i], rq[i] = C.uint8_t(p.Role), C.double(p.KVUsage), C.int32_t(p.Waiting), C.int32_t(p.Running)
	}
	e.mu.Unlock()
	if n == 0 {
		return errors.New("empty pool")
	}
	if rc := C.epp_pool_set(e.h, C.int32_t(n), &ids[0], &role[0], &kv[0], &wq[0], &rq[0], nil); rc != 0 {
		return lastErr(rc)
	}
	return nil
}

func (e *Engine) slotOf(name string) uint32 {
	if s, ok := e.slots[name]; ok {
		return s
	}
	var s uint32
	if n := len(e.free); n > 0 {
		s, e.free = e.free[n-1], e.free[:n-1]
	} else {
		s = uint32(len(e.slots))
	}
	e.slots[name] = s
	return s
}

// RemoveEndpoint mirrors indexer.RemovePod (indexer.go:167-182) / CleanUpInactivePods (plugin.go:99-122).
func (e *Engine) RemoveEndpoint(name string) {
	e.mu.Lock()
	s, ok := e.slots[name]
	if ok {
		delete(e.slots, name)
		e.free = append(e.free, s)
	}
	e.mu.Unlock()
	if ok {
		C.epp_index_remove_endpoint(e.h, C.uint32_t(s))
	}
}

// RetainEndpoints mirrors CleanUpInactivePods (approximateprefix/plugin.go:99-122): every indexed endpoint that is not
// in `active` loses its LRU and its postings (RemovePod), on the device, in one call.
func (e *Engine) RetainEndpoints(active []string) error {
	e.mu.Lock()
	ids := make([]C.uint32_t, 0, len(active))
	for _, name := range active {
		if s, ok := e.slots[name]; ok {
			ids = append(ids, C.uint32_t(s))
		}
	}
	e.mu.Unlock()
	var p *C.uint32_t
	if len(ids) > 0 {
		p = &ids[0]
	}
	if rc := C.epp_index_retain_endpoints(e.h, C.int32_t(len(ids)), p); rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// LoraState is fwkdl.Metrics.{ActiveModels, WaitingModels, MaxActiveModels} of one endpoint (metrics.go:26-42) with the
// adapter names already translated to model ids (epp_model_register).
type LoraState struct {
	Endpoint        string
	MaxActiveModels int
	Active, Waiting []uint32
}

// SetLora feeds the lora-affinity scorer (scorer/loraaffinity/lora_affinity.go:76-100); call it after SetPool of the
// same scrape.
func (e *Engine) SetLora(states []LoraState) error {
	n := len(states)
	if n == 0 {
		return nil
	}
	ids := make([]C.uint32_t, n)
	mx := make([]C.int32_t, n)
	cnt := make([]C.int32_t, n)
	var mep, mmo []C.uint32_t
	var mst []C.uint8_t
	e.mu.Lock()
	for i, st := range states {
		s := e.slotOf(st.Endpoint)
		ids[i], mx[i], cnt[i] = C.uint32_t(s), C.int32_t(st.MaxActiveModels), C.int32_t(len(st.Active)+len(st.Waiting))
		for _, m := range st.Active {
			mep, mmo, mst = append(mep, C.uint32_t(s)), append(mmo, C.uint32_t(m)), append(mst, 1)
		}
		for _, m := range st.Waiting {
			mep, mmo, mst = append(mep, C.uint32_t(s)), append(mmo, C.uint32_t(m)), append(mst, 2)
		}
	}
	e.mu.Unlock()
	var pe, pm *C.uint32_t
	var ps *C.uint8_t
	if len(mep) > 0 {
		pe, pm, ps = &mep[0], &mmo[0], &mst[0]
	}
	if rc := C.epp_pool_set_lora(e.h, C.int32_t(n), &ids[0], &mx[0], &cnt[0], C.int64_t(len(mep)), pe, pm, ps); rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// Decision is one epp_decision translated back to endpoint names.
type Decision struct {
	Err         error
	Pick        uint32
	Score       float64
	PrefillPick uint32 // EPP_NO_ENDPOINT when absent
	TotalBlocks int
	MatchBlocks int
}

// Schedule is what a Scheduler implementation (director.go:69-71) calls per request; requests from concurrent
// ext-proc goroutines (handlers/server.go:168) are coalesced by the micro-batcher into one epp_schedule call.
func (e *Engine) Schedule(ctx context.Context, modelID uint32, prompt []byte) (Decision, error) {
	return e.batch.submit(ctx, modelID, prompt)
}

// scheduleBatch is the single cgo crossing per batch.  Prompts are copied into a pinned C staging buffer
// (epp_host_alloc) so that no Go pointer is retained by C and the H2D copy is asynchronous.
func (e *Engine) scheduleBatch(stage unsafe.Pointer, offsets, lengths []C.uint64_t, modelIDs []C.uint32_t, out []C.epp_decision) error {
	var b C.epp_batch
	b.n_requests = C.int64_t(len(modelIDs))
	b.data = stage
	b.offsets = &offsets[0]
	b.lengths = &lengths[0]
	b.model_ids = &modelIDs[0]
	if rc := C.epp_schedule(e.h, &b, &out[0], nil, 1 /* keep_hashes: PreRequest follows */); rc != 0 {
		return lastErr(rc)
	}
	// PreRequest (approximateprefix/plugin.go:164-200): index the batch's hashes for the picked endpoints.
	if rc := C.epp_index_add_picked(e.h); rc != 0 {
		return lastErr(rc)
	}
	return nil
}

var errNoEndpoints = errors.New("no endpoints available for the given request") // scheduler_profile.go:119-121

// ---- endpoint-sharded pools (index larger than one GPU): exchanges over NVLink peer memory, include/epp_engine.h ----

// ShardSet restricts this replica's engine to the endpoint slots [begin, end); SetPool must follow.
func (e *Engine) ShardSet(begin, end uint32) error {
	if rc := C.epp_shard_set(e.h, C.uint32_t(begin), C.uint32_t(end)); rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// ShardP2PExport allocates this rank's exchange buffer and returns its 64-byte CUDA IPC handle; the replicas exchange
// the handles once over their control channel (any transport: they are plain bytes).
func (e *Engine) ShardP2PExport(maxRequests int) ([64]byte, error) {
	var h [64]byte
	var ptr C.uint64_t
	if rc := C.epp_shard_p2p_export(e.h, C.int64_t(maxRequests), (*C.uint8_t)(unsafe.Pointer(&h[0])), &ptr); rc != 0 {
		return h, lastErr(rc)
	}
	return h, nil
}

// ShardP2PConnect opens every peer's exchange buffer (handles[rank] is this rank's own and is ignored).
func (e *Engine) ShardP2PConnect(rank int, handles [][64]byte) error {
	flat := make([]byte, 0, 64*len(handles))
	for _, h := range handles {
		flat = append(flat, h[:]...)
	}
	if rc := C.epp_shard_p2p_connect(e.h, C.int32_t(len(handles)), C.int32_t(rank), unsafe.Pointer(&flat[0]), 1); rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// ShardScheduleP2P runs one batch of the sharded protocol; every replica calls it with the same device-resident batch
// (devPrompts / devOut are device pointers owned by the caller, e.g. filled by the tokenizer stage on the same GPU).
func (e *Engine) ShardScheduleP2P(devPrompts unsafe.Pointer, nRequests int, uniformLen uint64, devOut unsafe.Pointer) error {
	var b C.epp_batch
	b.n_requests = C.int64_t(nRequests)
	b.data = devPrompts
	b.uniform_len = C.uint64_t(uniformLen)
	b.flags = C.EPP_BATCH_DEVICE_PTRS
	if rc := C.epp_shard_schedule_p2p(e.h, &b, (*C.epp_decision)(devOut)); rc != 0 {
		return lastErr(rc)
	}
	return nil
}
