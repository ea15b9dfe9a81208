"""Host-side restatement of the index arithmetic of the dx-in-N kernel (conv_tc.cu: conv_dx_kernel,
dx_item, set_split) checked exhaustively on CPU: every output pixel of an image is produced by
exactly one (tile, block, row), every shared-memory row an MMA touches lies inside the halo tile
TMA delivered, and the split last round deals every tile exactly once."""
import itertools

import pytest

PITCH, STRIP, BLK = 66, 64, 126          # kPitch, kStrip, kDxBlk


def tile_rows(mb):                       # TileGeom<MB, CH>::kRows
    return 5 if mb == 1 else 7


def tile_geometry(t, mb):
    """(r0, base_flat) of tile t: first image row of the TMA box, tile-relative flat row of
    block 0 / dy = 0 (conv_dx_kernel, MMA warp)."""
    f0 = t * BLK * mb
    r0 = (f0 + PITCH - 1) // PITCH - 2
    return f0, r0, f0 - r0 * PITCH


@pytest.mark.parametrize("h,mb", [(64, 2), (64, 1), (40, 2), (7, 1), (2, 2), (256, 2), (23, 1)])
def test_every_pixel_once_and_operands_inside_the_halo_tile(h, mb):
    s_out = BLK * mb
    tiles = (h * PITCH + s_out - 1) // s_out
    seen = {}
    for t in range(tiles):
        f0, r0, base = tile_geometry(t, mb)
        assert 67 <= base <= 132
        for b in range(mb):
            # operand rows of the three window rows: smem flat = base + 126 b + dy*66 + r, r in [0,128)
            lo = base + BLK * b - PITCH
            hi = base + BLK * b + PITCH + 127
            assert lo >= 1 and hi < tile_rows(mb) * PITCH, (t, b, lo, hi)
            for r in range(1, BLK + 1):              # rows 0 and 127 belong to the neighbours
                f = f0 - 1 + BLK * b + r
                py, pc = divmod(f, PITCH)
                if pc < STRIP and py < h:
                    assert (py, pc) not in seen
                    seen[(py, pc)] = (t, b, r)
                    # the A row of tap (dy, dx) is image pixel (py+dy, pc+dx): smem column pc+dx+1
                    for dy, dx in itertools.product((-1, 0, 1), repeat=2):
                        a_row = base + BLK * b + dy * PITCH + (r + dx)      # block row r+dx, shift dy
                        srow, scol = divmod(a_row, PITCH)
                        assert (srow + r0, scol - 1) == (py + dy, pc + dx)
    assert len(seen) == h * STRIP


def dx_item(it, cta, grid, total, split_round, split_items, split_tile0):
    """conv_tc.cu: dx_item."""
    if split_round >= 0 and it >= split_round:
        if it > split_round or cta >= split_items:
            return None
        return split_tile0 + (cta >> 1), cta & 1
    tile = cta + it * grid
    return (tile, -1) if tile < total else None


def set_split(total, grid, mb, pays=True):
    """conv_tc.cu: set_split."""
    rounds, rem = divmod(total, grid)
    if mb == 2 and pays and rounds >= 1 and rem > 0 and 2 * rem <= grid:
        return rounds, 2 * rem, rounds * grid
    return -1, 0, 0


@pytest.mark.parametrize("total,grid", [(1088, 148), (51, 40), (34, 24), (34, 5), (148, 148), (149, 148), (200, 148),
                                         (10, 148), (296, 148), (1000, 7)])
def test_split_last_round_deals_every_block_once(total, grid):
    grid = min(grid, total)
    sr, si, st0 = set_split(total, grid, 2)
    work = {}
    most = 0
    for cta in range(grid):
        it, blocks = 0, 0
        while True:
            item = dx_item(it, cta, grid, total, sr, si, st0)
            if item is None:
                break
            tile, sel = item
            for b in ((0, 1) if sel < 0 else (sel,)):
                assert (tile, b) not in work
                work[(tile, b)] = cta
                blocks += 1
            if sel >= 0:      # a half-tile is always the CTA's last item (accumulator parities rely on it)
                assert dx_item(it + 1, cta, grid, total, sr, si, st0) is None
            it += 1
        most = max(most, blocks)
    assert len(work) == 2 * total
    if sr >= 0:               # the split saves half a round
        assert most == 2 * (total // grid) + 1


# ---------------------------------------------------------------------------------------------
# CTA pairs: which packed weight rows each CTA of a pair loads, where they sit in its shared memory,
# and which accumulator column they produce (conv_tc.cu: conv_pair_kernel / conv_dx_kernel<PAIR>).
# `tcgen05.mma.cta_group::2` takes the first N/2 rows of B from the even CTA and the rest from the
# odd one, both at the SAME shared-memory offset.

def pair_tap_layout(rank):
    """conv_pair_kernel weight producer: smem rows of one tap in CTA `rank` -> packed rows (tap-relative;
    packed rows 0-63 = W_hi, 64-127 = W_lo')."""
    x = [rank * 64 + r for r in range(64)]          # two 32-row boxes at +0 and +2048 B
    y = [rank * 32 + r for r in range(32)]          # one box at +4096 B
    return x, y


def test_conv5_pair_weight_split_reproduces_the_single_cta_columns():
    x0, y0 = pair_tap_layout(0)
    x1, y1 = pair_tap_layout(1)
    wide = x0 + x1              # B rows of the N=128 MMA in column order
    narrow = y0 + y1            # B rows of the N=64 MMA (written at column offset 64)
    # single-CTA kernel: columns 0-63 = W_hi couts, 64-127 = W_lo' couts; narrow = W_hi into 64-127
    assert wide == list(range(128))
    assert narrow == list(range(64))


def dx_pair_slab_layout(rank):
    """conv_dx_kernel<PAIR> weight producer: smem rows of one (chunk, dy) slab in CTA `rank` as
    (part, dx, cout) triples; boxes are {CH, 16 couts, 3 dx, 1 part} = 48 rows each."""
    def box(part, half):
        return [(part, dx, half * 16 + c) for dx in range(3) for c in range(16)]
    x = box(rank, 0) + box(rank, 1)                 # X: this CTA's 96 rows of the wide operand
    y = box(0, rank)                                # Y: W_hi half `rank` for the lo' phase
    return x, y


def test_dx_pair_columns_match_the_epilogue_mapping():
    x0, y0 = dx_pair_slab_layout(0)
    x1, y1 = dx_pair_slab_layout(1)
    wide = x0 + x1              # N = 192: columns 0-95 main (W_hi), 96-191 correction (W_lo')
    narrow = y0 + y1            # N = 96, accumulated at column offset 96
    for part in range(2):
        for dx in range(3):
            for cout in range(32):
                col = (cout // 16) * 48 + dx * 16 + cout % 16      # epilogue `drain` (PAIR branch)
                assert wide[part * 96 + col] == (part, dx, cout)
                if part == 0:
                    assert narrow[col] == (0, dx, cout)             # lo' x W_hi lands on the same channel
    assert len(wide) == 192 and len(set(wide)) == 192 and len(narrow) == 96


# ---------------------------------------------------------------------------------------------
# Epilogue store-transpose staging (conv_common.cuh: finish_slice32): shared-memory bank pressure of
# the shipped 80-byte pitch and of the experimental XOR-swizzled 64-byte pitch (-DBHSR_EPI_SWZ).
def _wavefronts(addresses_16B):
    """Shared-memory wavefronts of one 16-byte-per-lane access: 128-bit accesses are served a quarter
    warp (8 lanes, 128 B) at a time; a quarter needs as many passes as the busiest bank has distinct
    4-byte words (the model that matches ncu's bank-conflict counter for this epilogue)."""
    total = 0
    for q in range(0, len(addresses_16B), 8):
        per_bank = {}
        for a in addresses_16B[q:q + 8]:
            for wd in range(a // 4, a // 4 + 4):
                per_bank.setdefault(wd % 32, set()).add(wd)
        total += max(len(v) for v in per_bank.values())
    return total


def _staging(pitch, swizzle):
    def off(row, chunk):
        return row * pitch + (((chunk ^ ((row >> 1) & 3)) if swizzle else chunk) << 4)
    writes = max(_wavefronts([off(lane, g) for lane in range(32)]) for g in range(4))
    reads = max(_wavefronts([off(8 * j4 + (lane >> 2), lane & 3) for lane in range(32)]) for j4 in range(4))
    return writes, reads


def test_staging_layout_bank_pressure():
    # 512 bytes per instruction = 4 wavefronts at best
    assert _staging(80, False) == (4, 8)     # shipped: conflict-free writes, 2x on the reads (ncu: ~0.9 M conflicts / launch)
    assert _staging(64, True) == (4, 4)      # experimental layout (-DBHSR_EPI_SWZ): both conflict-free
    assert _staging(64, False) == (16, 4)    # an unswizzled 64-byte pitch would move the conflicts to the writes


# ---------------------------------------------------------------------------------------------
# Shared-memory plans of the launchers (conv_tc.cu: launch, launch_dx, launch_pair, launch_dx_pair)
# restated: every trunk layer fits the 227 KB limit and gets the ring depths DESIGN.md describes.
SMEM_LIMIT = 232448
K_MAX_A, K_MAX_W = 4, 32


def _tile_bytes(mb, ch):
    raw = (5 if mb == 1 else 7) * PITCH * ch * 2
    return raw, (raw + 1023) // 1024 * 1024


def plan_dx(cin, exact, mb, pair=False):
    ch = 32 if exact else 64
    npart = 2 if exact else 1
    n_chunks = (cin + ch - 1) // ch
    tail = (4 * K_MAX_A + 8 + 2 * K_MAX_W) * 8 + 16 + 2 * 64 * 4 + 64 + 8 * 32 * 80 + 2 * 2 * 4 * 64 * 4
    w_slab = (144 if pair else 96 * npart) * ch * 2
    a_stage = _tile_bytes(mb, ch)[1] * npart
    slabs = n_chunks * 3
    slots = lambda ns: max(0, (SMEM_LIMIT - 1024 - ns * a_stage - tail) // w_slab)
    if pair:
        ns, ws = 2, min(slots(2), K_MAX_W)
        resident = slabs <= ws
        ws = slabs if resident else ws
    else:
        ns = ws = None
        for cand in range(K_MAX_A, 1, -1):
            if slots(cand) >= slabs:
                ns, ws = cand, slots(cand)
                break
        if ns is None:
            for cand in range(K_MAX_A, 1, -1):
                if slots(cand) >= 6:
                    ns, ws = cand, slots(cand)
                    break
        ws = min(ws, K_MAX_W)
        resident = slabs <= ws
        ws = slabs if resident else ws
    total = 1024 + ns * a_stage + ws * w_slab + tail
    return dict(astages=ns, wslots=ws, resident=resident, bytes=total)


def test_dx_shared_memory_plans_of_the_trunk_layers():
    exp = {64: (True, 2), 96: (False, 2), 128: (False, 2), 160: (False, 2)}     # conv1..conv4, exact, MB=2
    for cin, (resident, ns) in exp.items():
        pl = plan_dx(cin, True, 2)
        assert pl["bytes"] <= SMEM_LIMIT and pl["resident"] == resident and pl["astages"] == ns, (cin, pl)
        assert pl["wslots"] >= 6 or pl["resident"]
    assert plan_dx(64, True, 1)["astages"] == 3 and plan_dx(64, True, 1)["resident"]       # MB=1: a third stage fits
    for cin in (64, 96, 128, 160):
        assert plan_dx(cin, False, 2)["bytes"] <= SMEM_LIMIT                                   # fast numerics
    # CTA pairs keep 144 of 192 weight rows: conv2's weights become resident as well
    assert plan_dx(96, True, 2, pair=True)["resident"] and not plan_dx(128, True, 2, pair=True)["resident"]
    assert all(plan_dx(c, True, 2, pair=True)["bytes"] <= SMEM_LIMIT for c in (64, 96, 128, 160))


def test_pair_kernel_shared_memory_plan():
    tail = (2 * K_MAX_A + 4 + 2 * K_MAX_W) * 8 + 16 + 2 * 64 * 4 + 64 + 4 * 32 * 80
    a_stage = _tile_bytes(2, 32)[1] * 2
    w_slab = 3 * 96 * 64
    wslots = (SMEM_LIMIT - 1024 - 2 * a_stage - tail) // w_slab
    assert wslots == 5 and 1024 + 2 * a_stage + wslots * w_slab + tail <= SMEM_LIMIT
    # bytes both CTAs report to the leader's barriers stay far below the mbarrier tx-count range (2^20 - 1)
    assert 2 * (2 * _tile_bytes(2, 32)[0]) < (1 << 20) and 2 * w_slab < (1 << 20)
