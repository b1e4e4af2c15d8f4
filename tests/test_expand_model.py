"""Host-side model of the neighbour-list expansion of csrc/mis_tilebuild.cuh (k_tile_expand): mask bits -> list entries in
ascending tile order, entry-parallel.  The CUDA kernel is checked entry for entry on the GPU (test_gpu_gather_modes.py); this
restates its ALGORITHM in numpy -- byte markers at the first entry of every non-empty mask byte, a running maximum over the
entries, a 256 x 8 select table, the transposed uint16 block layout of mis_tile.cuh -- against the obvious expansion, on random
and adversarial masks (empty, full, single bits at word edges), so that the index arithmetic is pinned without a device."""
import numpy as np
import pytest

TILE_G, TILE_BLOCK = 8, 64


def select_table():
    lut = np.zeros((256, 8), np.uint8)
    for v in range(256):
        pos = [b for b in range(8) if (v >> b) & 1]
        lut[v, :len(pos)] = pos
    return lut


def tile_list_pos(e):
    return (e & ~(TILE_BLOCK - 1)) | ((e % TILE_G) * 8) | ((e // TILE_G) & 7)


def expand_model(words):
    """words: uint32 mask words of one particle.  Returns (tile indices in list order, transposed uint16 blocks incl. pads = 0xffff)."""
    words = np.asarray(words, np.uint32)
    by = words.view(np.uint8)                                  # little-endian: byte u = bits 8 (u % 4) .. of word u // 4
    pop = np.array([bin(int(b)).count("1") for b in by])
    pre = np.concatenate([[0], np.cumsum(pop)[:-1]])           # first entry of byte u
    cnt = int(pop.sum())
    mk = np.zeros(max(cnt, 1), np.int64)
    for u in np.nonzero(pop)[0]:
        mk[pre[u]] = u + 1                                     # distinct positions: non-empty bytes have distinct first entries
    owner = np.maximum.accumulate(mk)[:cnt] - 1                # running max = the byte entry e belongs to
    lut = select_table()
    e = np.arange(cnt)
    t = owner * 8 + lut[by[owner], e - pre[owner]] if cnt else np.zeros(0, np.int64)
    nb = (cnt + TILE_BLOCK - 1) // TILE_BLOCK
    blocks = np.full(nb * TILE_BLOCK, 0xffff, np.uint16)
    for k in range(nb * TILE_BLOCK):
        blocks[tile_list_pos(k)] = (t[k] * 16) if k < cnt else 0xffff
    return t, blocks


def naive(words):
    bits = np.unpackbits(np.asarray(words, np.uint32).view(np.uint8), bitorder="little")
    return np.nonzero(bits)[0]


@pytest.mark.parametrize("seed", range(6))
def test_entry_parallel_expansion_equals_the_obvious_one(seed):
    rng = np.random.default_rng(seed)
    W = int(rng.integers(1, 129))
    dens = rng.choice([0.02, 0.15, 0.6])
    words = (rng.random((W, 32)) < dens).astype(np.uint32)
    # blobs: whole words set or cleared, as the masks of neighbouring records are
    words[rng.random(W) < 0.3] = 0
    words[rng.random(W) < 0.1] = 1
    packed = (words << np.arange(32, dtype=np.uint32)).sum(1).astype(np.uint32)
    t, blocks = expand_model(packed)
    ref = naive(packed)
    assert np.array_equal(t, ref)
    # the reader of mis_tile.cuh: lane g of a group loads the 16-byte word g of a block; in pass p it uses entry p of that word
    # = list entry p * G + g of the block
    for blk in range(len(blocks) // TILE_BLOCK):
        for g in range(TILE_G):
            for p in range(8):
                e = blk * TILE_BLOCK + p * TILE_G + g
                want = ref[e] * 16 if e < len(ref) else 0xffff
                assert blocks[blk * TILE_BLOCK + g * 8 + p] == want


@pytest.mark.parametrize("words", [[0], [0xffffffff], [1], [0x80000000], [0, 0x80000000, 1, 0], [0xffffffff] * 4, [0x01010101, 0x80808080]])
def test_edge_masks(words):
    t, _ = expand_model(np.array(words, np.uint32))
    assert np.array_equal(t, naive(np.array(words, np.uint32)))
