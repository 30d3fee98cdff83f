package eppcuda

/*
#include "epp_engine.h"
*/
import "C"

import (
	"context"
	"encoding/json"
	"errors"
	"unsafe"

	"github.com/llm-d/llm-d-router/pkg/epp/framework/interface/plugin"
	fwkrc "github.com/llm-d/llm-d-router/pkg/epp/framework/interface/requestcontrol"
	"github.com/llm-d/llm-d-router/pkg/epp/framework/interface/scheduling"
	attrprefix "github.com/llm-d/llm-d-router/pkg/epp/framework/plugins/datalayer/attribute/prefix"
)

// Plugin-parity mode: the engine behind the reference's per-plugin interfaces, one request per call (R = 1 batches
// through epp_prefix_match / epp_score).  Slower than Scheduler (no batching) -- it exists so that a deployment can
// replace ONE plugin (the prefix producer or a scorer) and keep the rest of its Go profile unchanged.
const (
	PrefixProducerType = "cuda-approx-prefix-cache-producer"
	WeightedScorerType = "cuda-weighted-scorer"
)

var shared *Engine // set by Register

// Register installs the factories (plugin.Register, registry.go:25-30).  Call it next to RegisterAllPlugins
// (cmd/epp/main.go:59) after the engine has been created.
func Register(e *Engine) {
	shared = e
	plugin.Register(PrefixProducerType, func(name string, _ json.RawMessage, _ plugin.Handle) (plugin.Plugin, error) {
		if shared == nil {
			return nil, errors.New("eppcuda: engine not created")
		}
		return &PrefixProducer{name: name, e: shared}, nil
	})
	plugin.Register(WeightedScorerType, func(name string, params json.RawMessage, _ plugin.Handle) (plugin.Plugin, error) {
		var p struct {
			Profile int `json:"profile"` // 0 primary, 1 prefill, 2 encode
			Scorer  int `json:"scorer"`  // index inside the profile, -1 = the ordered weighted sum
		}
		p.Scorer = -1
		if len(params) > 0 {
			if err := json.Unmarshal(params, &p); err != nil {
				return nil, err
			}
		}
		return &WeightedScorer{name: name, e: shared, profile: p.Profile, scorer: p.Scorer}, nil
	})
}

// matchRow runs epp_prefix_match for one request: dense matchBlocks per slot + totalBlocks.
func (e *Engine) matchRow(model uint32, prompt []byte) ([]int32, int32, error) {
	match := make([]int32, e.cfg.MaxEndpoints)
	var total C.int32_t
	var b C.epp_batch
	b.n_requests = 1
	b.uniform_len = C.uint64_t(len(prompt))
	if len(prompt) > 0 {
		b.data = unsafe.Pointer(&prompt[0]) // Go memory without inner pointers, for the duration of the call
	}
	// model_ids stays NULL for model 0; other models go through a one-element C array
	var mid *C.uint32_t
	if model != 0 {
		mid = (*C.uint32_t)(C.malloc(4))
		defer C.free(unsafe.Pointer(mid))
		*mid = C.uint32_t(model)
		b.model_ids = mid
	}
	if rc := C.epp_prefix_match(e.h, &b, (*C.int32_t)(unsafe.Pointer(&match[0])), &total); rc != C.EPP_OK {
		return nil, 0, lastError()
	}
	return match, int32(total), nil
}

// PrefixProducer implements requestcontrol.DataProducer (plugins.go:69-72): it leaves a *PrefixCacheMatchInfo under
// PrefixCacheMatchInfoKey on every endpoint, like approximateprefix.Produce (plugin.go:135-160).
type PrefixProducer struct {
	name string
	e    *Engine
}

var _ fwkrc.DataProducer = (*PrefixProducer)(nil)

func (p *PrefixProducer) TypedName() plugin.TypedName {
	return plugin.TypedName{Type: PrefixProducerType, Name: p.name}
}

func (p *PrefixProducer) Produces() map[string]any {
	return map[string]any{attrprefix.PrefixCacheMatchInfoKey: attrprefix.PrefixCacheMatchInfo{}}
}

func (p *PrefixProducer) Produce(_ context.Context, request *scheduling.InferenceRequest, pods []scheduling.Endpoint) error {
	prompt, err := UserInputBytes(request)
	if err != nil {
		return err
	}
	model, err := p.e.modelID(request.TargetModel, "")
	if err != nil {
		return err
	}
	match, total, err := p.e.matchRow(model, prompt)
	if err != nil {
		return err
	}
	p.e.mu.RLock()
	defer p.e.mu.RUnlock()
	for _, ep := range pods {
		m := 0
		if slot, ok := p.e.slotOf[ep.GetMetadata().NamespacedName]; ok {
			m = int(match[slot])
		}
		ep.Put(attrprefix.PrefixCacheMatchInfoKey, attrprefix.NewPrefixCacheMatchInfo(m, int(total), p.e.cfg.BlockSizeTokens))
	}
	return nil
}

// WeightedScorer implements scheduling.Scorer (plugins.go:68-72) with epp_score on the match info a producer attached.
type WeightedScorer struct {
	name            string
	e               *Engine
	profile, scorer int
}

var _ scheduling.Scorer = (*WeightedScorer)(nil)

func (s *WeightedScorer) TypedName() plugin.TypedName {
	return plugin.TypedName{Type: WeightedScorerType, Name: s.name}
}
func (s *WeightedScorer) Category() scheduling.ScorerCategory { return scheduling.Balance }

func (s *WeightedScorer) Score(_ context.Context, _ *scheduling.CycleState, _ *scheduling.InferenceRequest,
	pods []scheduling.Endpoint) map[scheduling.Endpoint]float64 {
	out := make(map[scheduling.Endpoint]float64, len(pods))
	E := s.e.cfg.MaxEndpoints
	match := make([]int32, E)
	total := int32(0)
	s.e.mu.RLock()
	for _, ep := range pods {
		if v, ok := ep.Get(attrprefix.PrefixCacheMatchInfoKey); ok {
			if info, ok := v.(*attrprefix.PrefixCacheMatchInfo); ok {
				if slot, ok := s.e.slotOf[ep.GetMetadata().NamespacedName]; ok {
					match[slot] = int32(info.MatchBlocks())
					total = int32(info.TotalBlocks())
				}
			}
		}
	}
	s.e.mu.RUnlock()
	scores := make([]float64, E)
	if rc := C.epp_score(s.e.h, 1, (*C.int32_t)(unsafe.Pointer(&match[0])), (*C.int32_t)(unsafe.Pointer(&total)), nil,
		C.int32_t(s.profile), C.int32_t(s.scorer), (*C.double)(unsafe.Pointer(&scores[0])), 0); rc != C.EPP_OK {
		return out // a scorer cannot fail in the reference's interface: missing endpoints contribute 0
	}
	s.e.mu.RLock()
	defer s.e.mu.RUnlock()
	for _, ep := range pods {
		if slot, ok := s.e.slotOf[ep.GetMetadata().NamespacedName]; ok && scores[slot] >= 0 {
			out[ep] = scores[slot]
		}
	}
	return out
}
