"""Pins oracle XXH64 / hashPrompt against tests/golden/xxh64_vectors.json (python-xxhash, independent
implementation of the public spec) and against the pure-Python restatement.  SURVEY App. A.1/A.2/B.2."""
import json
import os
import random
import struct

import pytest

from oracle import py_restatement as pr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "xxh64_vectors.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def test_xxh64_known_answers(orc):
    assert orc.xxh64(b"") == 0xEF46DB3751D8E999          # public spec vector
    assert orc.xxh64(b"a") == 0xD24EC4F1A98C6E5B
    assert orc.xxh64(b"test-model1") == 0x55B9CE9184DD8509
    assert orc.xxh64(b"synthetic-model") == 0xDEC636E2BE084E4C


def test_xxh64_raw_golden(orc, gold):
    for v in gold["raw"]:
        data = bytes.fromhex(v["hex"])
        want = int(v["xxh64"], 16)
        assert orc.xxh64(data) == want, len(data)
        assert pr.xxh64_pure(data) == want, len(data)


def test_hash_prompt_golden(orc, gold):
    for v in gold["prompts"]:
        data = bytes.fromhex(v["data_hex"])
        want = [int(h, 16) for h in v["hashes"]]
        got = orc.hash_prompt(data, v["model"].encode(), v["block_size_tokens"], v["max_blocks"], v["salt"].encode())
        assert got == want, v["name"]


def test_hash_prompt_survey_b2_vectors(orc):
    m = b"test-model1"
    assert orc.hash_prompt(b"aaaabbbb", m, 1, 256) == [0xCC8952E70E1C4504, 0x251707AC8B048154]
    assert orc.hash_prompt(b"aaaaaa", m, 1, 256) == [0xCC8952E70E1C4504, 0x857C9C463E7C65D4]
    assert orc.hash_prompt(b"aaaabbbbccccdddd", m, 1, 256) == [
        0xCC8952E70E1C4504, 0x251707AC8B048154, 0xA39267888AD6734A, 0x1BCEA087903CF2D3]
    assert orc.hash_prompt(b"aaaabbbb", m, 1, 256, b"s1") == [0xE01058411E46B07C, 0x3ED13228ACBA73F7]
    toks = struct.pack("<40I", *range(40))
    three = [0x8A787856B0C98B2D, 0xC57021EFCC0A8595, 0xA8A7E535C242C990]
    assert orc.hash_prompt(toks, b"synthetic-model", 16, 256) == three
    assert orc.hash_prompt(toks, b"synthetic-model", 16, 2) == three[:2]


def test_hash_prompt_reference_count_kats(orc):
    """Block-count KATs of approximateprefix/plugin_test.go (the reference pins counts, not values)."""
    m = b"test-model1"
    assert len(orc.hash_prompt(b"aaaabbbb", m, 1, 256)) == 2          # plugin_test.go:72
    assert len(orc.hash_prompt(b"aaaaaa", m, 1, 256)) == 2            # :187 (1 full + 1 partial)
    assert len(orc.hash_prompt(b"aaa", m, 1, 256)) == 0               # shorter than one block -> nil
    # :487-548  maxPrefixTokensToMatch=2, bs 1 token -> maxBlocks = 2 ; fallback maxPrefixBlocksToMatch=3
    assert len(orc.hash_prompt(b"aaaabbbbccccdddd", m, 1, 2)) == 2
    assert len(orc.hash_prompt(b"aaaabbbbccccdddd", m, 1, 3)) == 3
    # :440-485 autoTune: CacheBlockSize=16 -> 64-byte blocks: 128 chars = 2 blocks
    assert len(orc.hash_prompt(b"x" * 128, m, 16, 256)) == 2
    # prefix property: equal prefixes give equal leading hashes; first differing block changes all later ones
    a = orc.hash_prompt(b"aaaabbbbcccc", m, 1, 256)
    b = orc.hash_prompt(b"aaaabbbbdddd", m, 1, 256)
    assert a[:2] == b[:2] and a[2] != b[2]
    # different model -> different hashes (hashing.go:72-73)
    assert orc.hash_prompt(b"aaaabbbb", b"test-model2", 1, 256)[0] != a[0]


def test_hash_prompt_random_vs_python_restatement(orc):
    rng = random.Random(7)
    for _ in range(300):
        bs = rng.choice([1, 2, 3, 4, 7, 8, 16, 17, 31])
        n = rng.randint(0, bs * 4 * 12 + 9)
        data = bytes(rng.getrandbits(8) for _ in range(n))
        mb = rng.choice([1, 2, 5, 256])
        salt = rng.choice([b"", b"salty"])
        assert orc.hash_prompt(data, b"mdl", bs, mb, salt) == pr.hash_prompt(data, b"mdl", bs, mb, salt)
