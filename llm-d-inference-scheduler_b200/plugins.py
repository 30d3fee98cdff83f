"""Host-side mirror of the reference's plugin/scheduler interface FOR THIS PATH ONLY (same names, argument meaning
and error behaviour), so that parity tests read like the reference's own Go tests.  It is configuration and
marshalling: every score, match count and pick comes out of the CUDA engine through the C ABI.

Reference interfaces mirrored (paths relative to the reference root):
  fwksched.Endpoint / fwkdl.Metrics / EndpointMetadata   pkg/epp/framework/interface/scheduling/types.go:71-117
  Scorer / Filter / Picker / ProfileHandler               pkg/epp/framework/interface/scheduling/plugins.go:43-78
  NewWeightedScorer                                       pkg/epp/scheduling/weighted_scorer.go:24-40
  SchedulerProfile (WithFilters/WithScorers/WithPicker)   pkg/epp/scheduling/scheduler_profile.go
  Scheduler.Schedule                                      pkg/epp/scheduling/scheduler.go:54-102
  MaxScorePicker (maxNumOfEndpoints)                      .../picker/maxscore/picker.go:58-115
  prefix-cache-affinity-filter                            .../filter/prefixcacheaffinity/plugin.go:45-151
  approximateprefix dataProducer Produce / PreRequest     .../approximateprefix/plugin.go:135-200
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _capi as capi
from .engine import AffinityFilterSpec, Engine, ProfileSpec, ScorerSpec

RoleLabel = "llm-d.ai/role"
_ROLE_BY_LABEL = {
    "decode": capi.ROLE_DECODE, "prefill": capi.ROLE_PREFILL, "prefill-decode": capi.ROLE_PREFILL_DECODE,
    "both": capi.ROLE_BOTH, "encode": capi.ROLE_ENCODE, "encode-prefill": capi.ROLE_ENCODE_PREFILL,
    "encode-prefill-decode": capi.ROLE_ENCODE_PREFILL_DECODE,
}
RoleDecode, RolePrefill, RolePrefillDecode, RoleBoth = "decode", "prefill", "prefill-decode", "both"
RoleEncode, RoleEncodePrefill, RoleEncodePrefillDecode = "encode", "encode-prefill", "encode-prefill-decode"


class SchedulingError(RuntimeError):
    """Schedule returned an error (scheduler_profile.go:119-121 / disagg_profile_handler.go:335-338)."""


@dataclass
class Metrics:                                   # fwkdl.Metrics, framework/interface/datalayer/metrics.go:26-42
    WaitingQueueSize: int = 0
    KVCacheUsagePercent: float = 0.0
    RunningRequestsSize: int = 0
    CacheBlockSize: int = 0
    CacheNumBlocks: int = 0


@dataclass
class EndpointMetadata:
    Name: str
    Labels: dict = field(default_factory=dict)
    Address: str = ""


class Endpoint:
    """fwksched.Endpoint: metadata + metrics + attribute map (PrefixCacheMatchInfo lives under Put/Get)."""

    def __init__(self, metadata: EndpointMetadata, metrics: Metrics | None = None):
        self.metadata = metadata
        self.metrics = metrics or Metrics()
        self.attrs: dict = {}

    def GetMetadata(self):
        return self.metadata

    def GetMetrics(self):
        return self.metrics

    def Put(self, key, value):
        self.attrs[key] = value

    def Get(self, key):
        return self.attrs.get(key), key in self.attrs

    def role_code(self) -> int:
        if RoleLabel not in self.metadata.Labels:
            return capi.ROLE_NONE
        return _ROLE_BY_LABEL.get(self.metadata.Labels[RoleLabel], capi.ROLE_OTHER)


def NewEndpoint(metadata: EndpointMetadata, metrics: Metrics | None = None, attrs=None) -> Endpoint:
    return Endpoint(metadata, metrics)


PrefixCacheMatchInfoKey = "PrefixCacheMatchInfoKey"          # attribute/prefix/data_types.go:24


@dataclass
class PrefixCacheMatchInfo:                      # attribute/prefix/data_types.go:27-54
    matchBlocks: int
    totalBlocks: int
    blockSizeTokens: int

    def MatchBlocks(self):
        return self.matchBlocks

    def TotalBlocks(self):
        return self.totalBlocks

    def BlockSizeTokens(self):
        return self.blockSizeTokens


def NewPrefixCacheMatchInfo(match, total, block_size_tokens):
    return PrefixCacheMatchInfo(match, total, block_size_tokens)


# ---- scorers (configuration objects; the arithmetic is in pick_kernels.cu) ----
@dataclass
class _Scorer:
    kind: int
    param: float = 0.0
    type_name: str = ""
    column: int = 0
    param2: float = 0.0


def PrefixCacheScorer():                         # scorer/prefix "prefix-cache-scorer"
    return _Scorer(capi.SCORER_PREFIX, 0.0, "prefix-cache-scorer")


def KVCacheUtilizationScorer():                  # "kv-cache-utilization-scorer"
    return _Scorer(capi.SCORER_KV_UTIL, 0.0, "kv-cache-utilization-scorer")


def QueueScorer():                               # "queue-scorer"
    return _Scorer(capi.SCORER_QUEUE, 0.0, "queue-scorer")


def RunningRequestsScorer():                     # "running-requests-size-scorer"
    return _Scorer(capi.SCORER_RUNNING, 0.0, "running-requests-size-scorer")


def NewLoadAware(queueThreshold: int = 128):     # loadaware.NewLoadAware, load_aware.go:43-52
    return _Scorer(capi.SCORER_LOAD_AWARE, float(queueThreshold), "load-aware-scorer")


def TokenLoadScorer(column: int, queueThresholdTokens: int = 4194304):
    """ "token-load-scorer" (scorer/tokenload/token_load.go:84-112); `column` = ext column holding InFlightLoad.Tokens."""
    return _Scorer(capi.SCORER_TOKEN_LOAD, float(queueThresholdTokens), "token-load-scorer", column)


def NewActiveRequest(column: int, idleThreshold: int = 0, maxBusyScore: float = 1.0):
    """ "active-request-scorer" (scorer/activerequest/active_request.go:72-101, 140-173); `column` = ext column holding
    InFlightLoad.Requests."""
    return _Scorer(capi.SCORER_ACTIVE_REQUEST, float(maxBusyScore), "active-request-scorer", column, float(idleThreshold))


def LoraAffinityScorer():
    """ "lora-affinity-scorer" (scorer/loraaffinity/lora_affinity.go:76-100); residency via Engine.pool_set_lora."""
    return _Scorer(capi.SCORER_LORA_AFFINITY, 0.0, "lora-affinity-scorer")


def ExternalScorer(column: int):                 # host-computed column (e.g. lora-affinity)
    return _Scorer(capi.SCORER_EXTERNAL, float(column), "external")


@dataclass
class WeightedScorer:
    scorer: _Scorer
    weight: float


def NewWeightedScorer(scorer: _Scorer, weight: float) -> WeightedScorer:
    return WeightedScorer(scorer, float(weight))


@dataclass
class _RoleFilter:
    kind: int


def NewDecodeRole():
    return _RoleFilter(capi.FILTER_DECODE)


def NewPrefillRole():
    return _RoleFilter(capi.FILTER_PREFILL)


def NewEncodeRole():
    return _RoleFilter(capi.FILTER_ENCODE)


@dataclass
class PrefixCacheAffinityFilter:                 # filter/prefixcacheaffinity "prefix-cache-affinity-filter"; DefaultConfig :62-66
    affinityThreshold: float = 0.80
    explorationProbability: float = 0.01
    maxTTFTPenaltyMs: float = 5000.0
    ttft_column: int = -1                        # ext column carrying LatencyPredictionInfo.TTFT (-1: attribute absent)


@dataclass
class MaxScorePicker:                            # picker/maxscore; tie rule: include/epp_engine.h "Tie rule" / "Top-k"
    maxNumOfEndpoints: int = 1


def NewMaxScorePicker(maxNumOfEndpoints: int = 1):
    if maxNumOfEndpoints <= 0:                   # picker.go:59-61: invalid value -> DefaultMaxNumOfEndpoints
        maxNumOfEndpoints = 1
    return MaxScorePicker(maxNumOfEndpoints)


class SchedulerProfile:
    def __init__(self):
        self.filters: list = []
        self.scorers: list = []
        self.picker = MaxScorePicker()

    def WithFilters(self, *filters):
        roles = [f for f in filters if isinstance(f, _RoleFilter)]
        aff = [f for f in filters if isinstance(f, PrefixCacheAffinityFilter)]
        assert len(roles) <= 1 and len(aff) <= 1 and len(roles) + len(aff) == len(filters), \
            "a profile's filter chain is: [one role filter] [one prefix-cache-affinity-filter]"
        assert not (roles and aff) or filters.index(roles[0]) < filters.index(aff[0]), "the role filter comes first"
        self.filters = list(filters)
        return self

    def WithScorers(self, *scorers):
        self.scorers = list(scorers)
        return self

    def AddPlugins(self, *plugins):
        for p in plugins:
            if isinstance(p, WeightedScorer):
                self.scorers.append(p)
            elif isinstance(p, (_RoleFilter, PrefixCacheAffinityFilter)):
                self.filters.append(p)
            elif isinstance(p, MaxScorePicker):
                self.picker = p
        return None

    def WithPicker(self, picker):
        self.picker = picker
        return self

    def spec(self) -> ProfileSpec:
        roles = [f for f in self.filters if isinstance(f, _RoleFilter)]
        aff = [f for f in self.filters if isinstance(f, PrefixCacheAffinityFilter)]
        return ProfileSpec(roles[0].kind if roles else capi.FILTER_NONE,
                           [ScorerSpec(ws.scorer.kind, ws.weight, ws.scorer.param, ws.scorer.column, ws.scorer.param2)
                            for ws in self.scorers],
                           AffinityFilterSpec(aff[0].affinityThreshold, aff[0].explorationProbability,
                                              aff[0].maxTTFTPenaltyMs, aff[0].ttft_column) if aff else None)


def NewSchedulerProfile():
    return SchedulerProfile()


@dataclass
class PrefixBasedPDDecider:                      # disagg.NewPrefixBasedPDDecider(NonCachedTokens)
    NonCachedTokens: int


@dataclass
class AlwaysDisaggPDDecider:
    pass


@dataclass
class SingleProfileHandler:
    pass


@dataclass
class DisaggProfileHandler:                      # disagg.NewDisaggProfileHandler(decode, prefill, "", decider, nil)
    decodeProfile: str
    prefillProfile: str
    pdDecider: object = None


@dataclass
class ProfileRunResult:
    TargetEndpoints: list
    Score: float = 0.0
    TieCount: int = 1


@dataclass
class SchedulingResult:
    ProfileResults: dict
    PrimaryProfileName: str


@dataclass
class InferenceRequest:
    RequestID: str = ""
    TargetModel: str = ""
    Prompt: bytes = b""                          # Body.Completions.Prompt.Raw as bytes
    CacheSalt: str = ""


class Scheduler:
    """scheduling.NewSchedulerWithConfig(NewSchedulerConfig(profileHandler, profiles)) over the CUDA engine.

    Schedule() takes a LIST of requests (a batch against one frozen snapshot) and the candidate
    endpoints; PrefixCacheMatchInfo already Put on the endpoints is honoured exactly like the reference's scheduler
    tests do (disagg/scheduler_test.go:264-268); otherwise pass match info via `produce`."""

    def __init__(self, profileHandler, profiles: dict, *, max_endpoints: int = 64, block_size_tokens: int = 16,
                 max_prefix_blocks: int = 256, ext_columns: int = 0, device: int = 0, tie_seed: int = 0):
        self.handler = profileHandler
        self.profiles = profiles
        if isinstance(profileHandler, DisaggProfileHandler):
            primary = profiles[profileHandler.decodeProfile].spec()
            prefill = profiles[profileHandler.prefillProfile].spec() if profileHandler.prefillProfile in profiles else None
            self.primary_name = profileHandler.decodeProfile
            self.prefill_name = profileHandler.prefillProfile
            dec = profileHandler.pdDecider
            nct = dec.NonCachedTokens if isinstance(dec, PrefixBasedPDDecider) else 0
            always = isinstance(dec, AlwaysDisaggPDDecider)
            if prefill is not None and dec is None:
                prefill = None               # no decider => the prefill stage never runs (disagg_profile_handler.go:298)
        else:
            (self.primary_name, prof), = profiles.items()
            primary, prefill, nct, always = prof.spec(), None, 0, False
            self.prefill_name = None
        ks = {p.picker.maxNumOfEndpoints for p in profiles.values()}
        assert len(ks) == 1, "the engine has one maxNumOfEndpoints for all profiles"
        self.pick_k = ks.pop()
        self.engine = Engine(max_endpoints, primary, prefill, device=device, block_size_tokens=block_size_tokens,
                             max_prefix_blocks=max_prefix_blocks, non_cached_tokens=nct, always_disagg=always,
                             n_ext_cols=ext_columns, tie_seed=tie_seed, pick_k=self.pick_k)
        self.block_size_tokens = block_size_tokens

    def _push_pool(self, endpoints, ext=None):
        n = len(endpoints)
        self.engine.pool_set(np.arange(n, dtype=np.uint32), [e.role_code() for e in endpoints],
                             [e.metrics.KVCacheUsagePercent for e in endpoints],
                             [e.metrics.WaitingQueueSize for e in endpoints],
                             [e.metrics.RunningRequestsSize for e in endpoints], ext)

    def Schedule(self, requests, candidateEndpoints, ext=None):
        single = isinstance(requests, InferenceRequest)
        reqs = [requests] if single else list(requests)
        eps = list(candidateEndpoints)
        if not eps:
            raise SchedulingError("no endpoints available for the given request")
        self._push_pool(eps, ext)
        E = self.engine.E
        match = np.zeros((len(reqs), E), dtype=np.int32)
        total = np.zeros(len(reqs), dtype=np.int32)
        bst = self.block_size_tokens
        for i, ep in enumerate(eps):                     # injected PrefixCacheMatchInfo (same for every request)
            info, ok = ep.Get(PrefixCacheMatchInfoKey)
            if ok and info is not None:
                match[:, i] = info.MatchBlocks()
                total[:] = info.TotalBlocks()
                bst = info.BlockSizeTokens()
        in_len = np.array([len(r.Prompt) for r in reqs], dtype=np.int64)
        if self.pick_k > 1:
            dec, det, lists = self.engine.schedule_with_match(match, total, in_len, bst, topk=True)
        else:
            dec, det = self.engine.schedule_with_match(match, total, in_len, bst)
            lists = {"primary": dec["pick"].reshape(-1, 1), "prefill": dec["prefill_pick"].reshape(-1, 1)}
        out = [self._result(dec[i], det[i], eps, lists["primary"][i], lists["prefill"][i]) for i in range(len(reqs))]
        return out[0] if single else out

    def _result(self, d, dd, eps, primary, prefill):
        if d["status"] != 0:
            raise SchedulingError("failed to find available decode workers" if self.prefill_name
                                  else "no endpoints available for the given request")
        targets = lambda row: [eps[int(s)] for s in row if int(s) != capi.EPP_NO_ENDPOINT]     # picker order, k at most
        res = {self.primary_name: ProfileRunResult(targets(primary), float(d["score"]), int(d["tie_count"]))}
        if self.prefill_name and d["prefill_pick"] != capi.EPP_NO_ENDPOINT:
            res[self.prefill_name] = ProfileRunResult(targets(prefill), float(dd["prefill_score"]),
                                                      int(dd["prefill_tie_count"]))
        return SchedulingResult(res, self.primary_name)


class ApproxPrefixCacheProducer:
    """approximateprefix dataProducer: Produce (plugin.go:135-160) + PreRequest (plugin.go:164-200) on the engine."""

    def __init__(self, engine: Engine, model: bytes):
        self.engine = engine
        self.model_id = engine.register_model(model)
        self._last = None

    def Produce(self, prompts: list, endpoints: list):
        data = np.frombuffer(b"".join(prompts), dtype=np.uint8) if prompts else np.zeros(0, np.uint8)
        offs = np.zeros(len(prompts) + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in prompts], out=offs[1:])
        if data.size == 0:
            data = np.zeros(16, np.uint8)
        match, total = self.engine.prefix_match(data, offsets=offs,
                                                model_ids=np.full(len(prompts), self.model_id, np.uint32))
        hashes, nb = self.engine.hash_prompts(data, offsets=offs,
                                              model_ids=np.full(len(prompts), self.model_id, np.uint32))
        self._last = (hashes, nb)
        bst = self.engine.cfg.block_size_tokens
        if len(prompts) == 1:
            for i, ep in enumerate(endpoints):
                ep.Put(PrefixCacheMatchInfoKey, NewPrefixCacheMatchInfo(int(match[0, i]), int(total[0]), bst))
        return match, total

    def PreRequest(self, request_index: int, target_slots: list):
        hashes, nb = self._last
        for s in target_slots:
            self.engine.index_add(s, hashes[request_index, : nb[request_index]])
