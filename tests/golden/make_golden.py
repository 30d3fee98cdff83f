"""Generates tests/golden/xxh64_vectors.json with python-xxhash 3.7.0 (an implementation of the public
XXH64 spec that is independent of both oracle/epp_oracle.c and the CUDA kernels), following SURVEY.md
App. A.1/A.2.  The reference's own tests pin no hash VALUES (only counts/equalities), so these vectors
are what pins them.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import struct

import xxhash

HERE = os.path.dirname(os.path.abspath(__file__))


def x64(b: bytes) -> int:
    return xxhash.xxh64(b, seed=0).intdigest()


def hash_prompt(data: bytes, model: bytes, bs_tokens: int, max_blocks: int, salt: bytes = b""):
    bs = bs_tokens * 4
    if bs <= 0 or len(data) < bs:
        return []
    if len(data) > bs * max_blocks:
        data = data[: bs * max_blocks]
    prev = x64(model + salt)
    out = []
    i = 0
    while i + bs <= len(data):
        prev = x64(data[i:i + bs] + struct.pack("<Q", prev))
        out.append(prev)
        i += bs
    if i < len(data):
        prev = x64(data[i:] + struct.pack("<Q", prev))
        out.append(prev)
    return out


def main():
    rng = random.Random(0xE99)
    raw = []
    for n in list(range(0, 100)) + [127, 128, 129, 255, 256, 1000, 4096, 16384 + 8]:
        b = bytes(rng.getrandbits(8) for _ in range(n))
        raw.append({"hex": b.hex(), "xxh64": f"{x64(b):016x}"})
    raw.append({"hex": "", "xxh64": f"{x64(b''):016x}"})
    raw.append({"hex": b"a".hex(), "xxh64": f"{x64(b'a'):016x}"})

    prompts = []

    def add(name, data, model, bs_tokens, max_blocks, salt=b""):
        prompts.append({"name": name, "data_hex": data.hex(), "model": model.decode(), "salt": salt.decode(),
                        "block_size_tokens": bs_tokens, "max_blocks": max_blocks,
                        "hashes": [f"{h:016x}" for h in hash_prompt(data, model, bs_tokens, max_blocks, salt)]})

    # SURVEY App. B.2 rows
    add("aaaabbbb_bs1", b"aaaabbbb", b"test-model1", 1, 256)
    add("aaaaaa_partial", b"aaaaaa", b"test-model1", 1, 256)
    add("abcd16", b"aaaabbbbccccdddd", b"test-model1", 1, 256)
    add("aaaabbbb_salt", b"aaaabbbb", b"test-model1", 1, 256, b"s1")
    toks = struct.pack("<40I", *range(40))
    add("tokens0_39_bs16", toks, b"synthetic-model", 16, 256)
    add("tokens0_39_bs16_max2", toks, b"synthetic-model", 16, 2)
    add("too_short", b"abc", b"m", 1, 256)
    # random prompts: every (block size, tail length) class incl. bs not multiple of 32/8
    for bs_tokens in (1, 2, 3, 5, 8, 9, 16, 32, 33, 64):
        for extra in (0, 1, 3, 4, 7, 8, 13):
            nblk = rng.randint(1, 9)
            n = bs_tokens * 4 * nblk + extra
            data = bytes(rng.getrandbits(8) for _ in range(n))
            add(f"rand_bs{bs_tokens}_x{extra}", data, b"synthetic-model", bs_tokens, rng.choice([2, 4, 256]))
    # one full-size BASELINE config-3 prompt (4096 tokens -> 256 blocks) and a config-4 one (8192 -> 512)
    for T, mb in ((4096, 256), (8192, 512), (5000, 256)):
        data = struct.pack(f"<{T}I", *[rng.randrange(128000) for _ in range(T)])
        add(f"tokens_T{T}", data, b"synthetic-model", 16, mb)

    with open(os.path.join(HERE, "xxh64_vectors.json"), "w") as f:
        json.dump({"generator": "python-xxhash " + xxhash.VERSION, "raw": raw, "prompts": prompts}, f)
    print("wrote", len(raw), "raw +", len(prompts), "prompt vectors")


if __name__ == "__main__":
    main()
