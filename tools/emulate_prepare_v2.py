"""Thread-by-thread Python transliteration of the index logic of prepare_weights_kernel and of the experimental
prepare_weights_v2_kernel (csrc/prepare.cu) on random job tables with ragged shapes: both must write every destination
element exactly once, with the same value.  A desk check of the tile / group / transpose algebra written while no GPU
was available; it says nothing about barriers or shared-memory hazards -- the GPU test for that is
tests/test_gpu_training.py::test_inplace_weight_refresh_equals_rebuild under W2V2_PREP_V2=1."""
import numpy as np
rng = np.random.default_rng(0)

def make_jobs(n):
    jobs, tile = [], 0
    for i in range(n):
        R, C = int(rng.integers(1, 100)), int(rng.integers(1, 100))
        has16, hasT, has32 = rng.random() < 0.8, rng.random() < 0.6, rng.random() < 0.3
        if not (has16 or hasT or has32): has16 = True
        ld = C + int(rng.integers(0, 9)); ldt = R + int(rng.integers(0, 9))
        j = dict(src=rng.normal(size=(R, C)).astype(np.float32), R=R, C=C, ld=ld, ldt=ldt,
                 scale=np.float32(rng.choice([1.0, 0.125])), scale_t=np.float32(rng.choice([1.0, 0.125])), tile_begin=tile,
                 d16=np.full((R, ld), np.nan, np.float32) if has16 else None,
                 d32=np.full((R, ld), np.nan, np.float32) if has32 else None,
                 dT=np.full((C, ldt), np.nan, np.float32) if hasT else None,
                 w16=np.zeros((R, ld), int) if has16 else None, wT=np.zeros((C, ldt), int) if hasT else None)
        tile += ((R + 31) // 32) * ((C + 31) // 32)
        jobs.append(j)
    return jobs, tile

def find(jobs, t):
    lo, hi = 0, len(jobs) - 1
    while lo < hi:
        mid = (lo + hi + 1) >> 1
        if jobs[mid]["tile_begin"] <= t: lo = mid
        else: hi = mid - 1
    return lo

def h(x): return np.float32(np.float16(x))

def v1(jobs, total, grid):
    for b in range(grid):
        for t in range(b, total, grid):
            j = jobs[find(jobs, t)]
            tiles_c = (j["C"] + 31) // 32
            local = t - j["tile_begin"]
            r0, c0 = (local // tiles_c) * 32, (local % tiles_c) * 32
            tile = np.zeros((32, 33), np.float32)
            for ty in range(8):
                for tx in range(32):
                    for i in range(ty, 32, 8):
                        r, c = r0 + i, c0 + tx
                        v = np.float32(0)
                        if r < j["R"] and c < j["C"]:
                            raw = j["src"][r, c]
                            v = raw * j["scale"]
                            if j["d16"] is not None: j["d16"][r, c] = h(v); j["w16"][r, c] += 1
                            if j["d32"] is not None: j["d32"][r, c] = v
                            v = raw * j["scale_t"]
                        tile[i, tx] = v
            if j["dT"] is None: continue
            for ty in range(8):
                for tx in range(32):
                    for i in range(ty, 32, 8):
                        c, r = c0 + i, r0 + tx
                        if c < j["C"] and r < j["R"]: j["dT"][c, r] = h(tile[tx, i]); j["wT"][c, r] += 1

def v2(jobs, total, grid, G=4):
    for b in range(grid):
        t0 = b * G
        while t0 < total:
            sjob = [find(jobs, t0 + g) if t0 + g < total else -1 for g in range(G)]
            raw = np.zeros((8, 32, G, 4), np.float32)
            for ty in range(8):
                for tx in range(32):
                    for g in range(G):
                        if sjob[g] < 0: continue
                        j = jobs[sjob[g]]
                        tiles_c = (j["C"] + 31) // 32
                        local = t0 + g - j["tile_begin"]
                        r0, c = (local // tiles_c) * 32, (local % tiles_c) * 32 + tx
                        for k in range(4):
                            r = r0 + ty + 8 * k
                            if r < j["R"] and c < j["C"]: raw[ty, tx, g, k] = j["src"][r, c]
            tile = np.zeros((G, 32, 33), np.float32)
            any_t = False
            for ty in range(8):
                for tx in range(32):
                    for g in range(G):
                        if sjob[g] < 0: continue
                        j = jobs[sjob[g]]
                        tiles_c = (j["C"] + 31) // 32
                        local = t0 + g - j["tile_begin"]
                        r0, c = (local // tiles_c) * 32, (local % tiles_c) * 32 + tx
                        any_t |= j["dT"] is not None
                        for k in range(4):
                            i = ty + 8 * k; r = r0 + i
                            vt = np.float32(0)
                            if r < j["R"] and c < j["C"]:
                                v = raw[ty, tx, g, k] * j["scale"]
                                if j["d16"] is not None: j["d16"][r, c] = h(v); j["w16"][r, c] += 1
                                if j["d32"] is not None: j["d32"][r, c] = v
                                vt = raw[ty, tx, g, k] * j["scale_t"]
                            tile[g, i, tx] = vt
            if any_t:
                for ty in range(8):
                    for tx in range(32):
                        for g in range(G):
                            if sjob[g] < 0: continue
                            j = jobs[sjob[g]]
                            if j["dT"] is None: continue
                            tiles_c = (j["C"] + 31) // 32
                            local = t0 + g - j["tile_begin"]
                            r0, c0 = (local // tiles_c) * 32, (local % tiles_c) * 32
                            for k in range(4):
                                i = ty + 8 * k
                                c, r = c0 + i, r0 + tx
                                if c < j["C"] and r < j["R"]: j["dT"][c, r] = h(tile[g, tx, i]); j["wT"][c, r] += 1
            t0 += grid * G

import copy
for trial in range(3):
    jobs, total = make_jobs(7)
    a, b = copy.deepcopy(jobs), copy.deepcopy(jobs)
    v1(a, total, grid=5); v2(b, total, grid=3)
    for ja, jb in zip(a, b):
        for k in ("d16", "d32", "dT"):
            if ja[k] is not None:
                assert np.array_equal(ja[k], jb[k], equal_nan=True), (trial, k)
        if ja["d16"] is not None:
            assert (jb["w16"][:, :ja["C"]] == 1).all() and (ja["w16"][:, :ja["C"]] == 1).all()
            assert np.array_equal(jb["d16"][:, :ja["C"]], ja["src"].astype(np.float32) * 0 + np.vectorize(h)(ja["src"] * ja["scale"]))
        if ja["dT"] is not None:
            assert (jb["wT"][:, :ja["R"]] == 1).all()
            assert np.array_equal(jb["dT"][:, :ja["R"]], np.vectorize(h)(ja["src"] * ja["scale_t"]).T)
    print("trial", trial, "tiles", total, "ok")
