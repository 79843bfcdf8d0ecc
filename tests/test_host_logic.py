"""Host logic of the library that needs no device: the split-K plan of the persistent dequant-GEMM (csrc/gemm_umma.cuh
umma2_plan_split: longest chain of K steps per CTA + the reduce launch) at the shapes batched decode runs on a 148-SM B200."""
import pytest


def _chain(tiles, blocks, k_split, n_sm=148):
    """K steps (64 elements each) of the CTA with the most work, items dealt round-robin."""
    if not k_split:
        return -(-tiles // n_sm) * blocks * 4
    kb = k_split // 256
    n_z = -(-blocks // kb)
    return -(-(tiles * n_z) // n_sm) * kb * 4


@pytest.mark.parametrize("name,n_rows,K,want", [("q / o", 4096, 4096, 1024), ("k / v", 1024, 4096, 512), ("gate / up", 14336, 4096, 0),
                                                ("down", 4096, 14336, 3584), ("vocab head", 128256, 4096, 0)])
def test_split_plan_llama3_8b_batch_32(b200, name, n_rows, K, want):
    got = b200.plan_split(n_rows, K, 32)
    assert got == want, name
    tiles, blocks = -(-n_rows // 128), K // 256
    # never a longer chain than the unsplit launch, and a split must pay for its reduce launch (8 steps)
    assert _chain(tiles, blocks, got) + (8 if got else 0) <= _chain(tiles, blocks, 0)
    for s in range(2, 9):
        ks = -(-blocks // s) * 256
        if ks < K and s <= blocks // 2:
            assert _chain(tiles, blocks, got) + (8 if got else 0) <= _chain(tiles, blocks, ks) + 8


def test_round1_rule_split_gate_up_in_three(b200):
    """What the persistent kernel's plan replaced: 3 K ranges = 336 items = 3 rounds of 24 steps on 148 CTAs, against 64 unsplit."""
    ks = b200.plan_split(14336, 4096, 32, persistent=False)
    assert ks == 1536 and _chain(112, 16, ks) == 72 > _chain(112, 16, 0) == 64


def test_split_plan_edges(b200):
    assert b200.plan_split(4096, 4096, 65) == 0                      # wide token tiles are never split
    assert b200.plan_split(4096, 256, 8) == 0 and b200.plan_split(4096, 512, 8) == 0   # nothing to split
    assert b200.plan_split(128, 8192, 8, n_sm=4) == 2048             # one tile, four SMs: four ranges of eight blocks (eight ranges tie)
    with pytest.raises(b200.InvalidArgument):
        b200.plan_split(4096, 4000, 8)
