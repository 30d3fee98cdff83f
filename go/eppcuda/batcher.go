package eppcuda

/*
#include "epp_engine.h"
*/
import "C"

import (
	"context"
	"time"
	"unsafe"
)

// batcher coalesces per-request Schedule calls into one epp_schedule batch (max size / max delay), the only
// piece of host logic the synchronous per-request Go plugin API needs on top of the engine (SURVEY.md section 7,
// "Batching inside a request/response server").  All requests of a batch see ONE frozen snapshot (App. A.8).
type pending struct {
	model  uint32
	prompt []byte
	done   chan Decision
}

type batcher struct {
	e        *Engine
	in       chan pending
	maxBatch int
	maxDelay time.Duration
	stage    unsafe.Pointer // pinned, epp_host_alloc
	stageCap int
}

func newBatcher(e *Engine, maxBatch int, maxDelay time.Duration) *batcher {
	if maxBatch <= 0 {
		maxBatch = 4096
	}
	if maxDelay <= 0 {
		maxDelay = 200 * time.Microsecond
	}
	b := &batcher{e: e, in: make(chan pending, 4*maxBatch), maxBatch: maxBatch, maxDelay: maxDelay}
	go b.loop()
	return b
}

func (b *batcher) submit(ctx context.Context, model uint32, prompt []byte) (Decision, error) {
	p := pending{model: model, prompt: prompt, done: make(chan Decision, 1)}
	select {
	case b.in <- p:
	case <-ctx.Done():
		return Decision{}, ctx.Err()
	}
	select {
	case d := <-p.done:
		return d, d.Err
	case <-ctx.Done(): // same contract as the 400 ms producer timeout (director.go:55): the caller proceeds without us
		return Decision{}, ctx.Err()
	}
}

func (b *batcher) loop() {
	for first := range b.in {
		batch := []pending{first}
		timer := time.NewTimer(b.maxDelay)
	fill:
		for len(batch) < b.maxBatch {
			select {
			case p := <-b.in:
				batch = append(batch, p)
			case <-timer.C:
				break fill
			}
		}
		timer.Stop()
		b.flush(batch)
	}
}

func (b *batcher) flush(batch []pending) {
	// Prompts are laid out back to back with every START padded to a 32-byte boundary (selects the 256-bit-load
	// kernel); true lengths travel in epp_batch.lengths.
	offsets := make([]C.uint64_t, len(batch)+1)
	lengths := make([]C.uint64_t, len(batch))
	ids := make([]C.uint32_t, len(batch))
	total := 0
	for i, p := range batch {
		offsets[i] = C.uint64_t(total)
		lengths[i] = C.uint64_t(len(p.prompt))
		ids[i] = C.uint32_t(p.model)
		total += (len(p.prompt) + 31) &^ 31
	}
	offsets[len(batch)] = C.uint64_t(total)
	if total > b.stageCap {
		if b.stage != nil {
			C.epp_host_free(b.stage)
		}
		C.epp_host_alloc(C.size_t(total+64), &b.stage) // pinned C memory: cgo may keep it, H2D is asynchronous
		b.stageCap = total
	}
	for i, p := range batch {
		copy(unsafe.Slice((*byte)(unsafe.Add(b.stage, int(offsets[i]))), len(p.prompt)), p.prompt)
	}
	out := make([]C.epp_decision, len(batch))
	err := b.e.scheduleBatch(b.stage, offsets, lengths, ids, out)
	for i, p := range batch {
		d := Decision{Err: err}
		if err == nil {
			o := out[i]
			if o.status != 0 {
				d.Err = errNoEndpoints // -> ResourceExhausted / 429 (director.go:245)
			}
			d.Pick, d.Score, d.PrefillPick = uint32(o.pick), float64(o.score), uint32(o.prefill_pick)
			d.TotalBlocks, d.MatchBlocks = int(o.total_blocks), int(o.match_blocks)
		}
		p.done <- d
	}
}
