import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import epp_b200 as epp
from oracle import pyoracle as orc
E = int(os.environ.get("DBG_E", "12"))
prompt = bytes(range(200)) * 3     # 600 bytes -> 9 full blocks + partial
h = orc.hash_prompt(prompt, b"mdl", 16, 20)
blob = np.zeros(1024, np.uint8); blob[:len(prompt)] = np.frombuffer(prompt, np.uint8)
for label, pairs in (("empty", ([], [])), ("ep2 x3", (h[:3], [2, 2, 2])), ("ep0 x2", (h[:2], [0, 0])), ("ep5 x1", (h[:1], [5]))):
    eng = epp.Engine(E, epp.ProfileSpec(0, [epp.ScorerSpec(2, 2.0), epp.ScorerSpec(1, 2.0), epp.ScorerSpec(0, 3.0)]), max_prefix_blocks=20)
    eng.register_model(b"mdl")
    kv = np.linspace(0, 0.9, E); waiting = np.arange(E, dtype=np.int32) % 3
    eng.pool_set(np.arange(E), np.zeros(E, np.uint8), kv, waiting)
    eng.index_load_snapshot(np.array(pairs[0], np.uint64), np.array(pairs[1], np.uint32))
    dec, det = eng.schedule(blob, offsets=np.array([0, 608], np.uint64), lengths=np.array([600], np.uint64))
    m, t = eng.prefix_match(blob, offsets=np.array([0, 608], np.uint64), lengths=np.array([600], np.uint64))
    sc = eng.score(m, t)
    print(label, 'dec', dec[0], 'best by dense score:', int(np.argmax(sc[0])), float(sc[0].max()))
    eng.close()
