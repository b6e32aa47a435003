"""CPU-side checks of the C-ABI boundary: the library loads and exports exactly what
include/yc_b200.h declares; descriptor structs match the header layout; host-side logic."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "yc_b200.h")).read()
    return sorted(set(re.findall(r"YC_API\s+[\w\s\*]+?\b(yc_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from yolo_continuous_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(_lib.lib, n), n
    assert sorted(_lib.EXPORTS) == names
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == names  # nothing else leaks out of the shared object
    assert _lib.lib.yc_version() == 100


def test_struct_layout_matches_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    from yolo_continuous_b200 import _lib
    src = tmp_path / "probe.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "yc_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(yc_head_level), sizeof(yc_head_desc), sizeof(yc_nms_params),
         offsetof(yc_head_level, anchor_wh), offsetof(yc_head_desc, level), offsetof(yc_head_desc, z),
         offsetof(yc_head_desc, bins), offsetof(yc_nms_params, nms_thres), offsetof(yc_nms_params, image_hw));
  return 0; }''')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(_lib.HeadLevel), C.sizeof(_lib.HeadDesc), C.sizeof(_lib.NmsParams),
            _lib.HeadLevel.anchor_wh.offset, _lib.HeadDesc.level.offset, _lib.HeadDesc.z.offset,
            _lib.HeadDesc.bins.offset, _lib.NmsParams.nms_thres.offset, _lib.NmsParams.image_hw.offset]
    assert got == want


def test_size_queries_need_no_gpu():
    from yolo_continuous_b200 import _lib
    assert _lib.lib.yc_head_pack_bytes(255, 256) > 255 * 256 * 4
    assert _lib.lib.yc_head_pack_bytes(0, 256) == 0
    assert _lib.lib.yc_nms_workspace_bytes(64, 25200, 80) > 64 * 25200 * 60
    assert _lib.lib.yc_nms_workspace_bytes(0, 1, 1) == 0


def test_product_refuses_cpu_tensors():
    import torch
    from yolo_continuous_b200 import _lib, detect
    from yolo_continuous_b200.nets import IDetect
    from yolo_continuous_b200.utils import bbox
    with pytest.raises(_lib.YcError):
        detect.non_max_suppression(torch.zeros(1, 8, 85), 80, (640, 640), (640, 640), True)
    with pytest.raises(_lib.YcError):
        bbox.box_iou(torch.zeros(2, 4), torch.zeros(3, 4))
    head = IDetect(2, [[10, 13, 16, 30, 33, 23]], (8,)).eval()
    head.stride = torch.tensor([8.0])
    with pytest.raises(_lib.YcError):
        head([torch.zeros(1, 8, 4, 4)])


def test_state_dict_keys_match_reference_fixtures():
    """Checkpoint compatibility: same keys and shapes as the reference heads."""
    import torch
    from helpers import load
    from yolo_continuous_b200.nets import IAuxDetect, IBin, IDetect
    coco = [[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]]
    for cls, name, ch in ((IDetect, "idetect_nc80", (16, 32, 64)), (IAuxDetect, "iaux_nc80", (16, 32, 64) * 2),
                          (IBin, "ibin_nc80", (16, 32, 64))):
        fx = load(name)
        ref = {k[4:].replace("__", "."): v.shape for k, v in fx.items() if k.startswith("sd__")}
        head = cls(80, coco, ch)
        mine = {k: tuple(v.shape) for k, v in head.state_dict().items()}
        assert mine == {k: tuple(s) for k, s in ref.items()}, name
        head.load_state_dict({k[4:].replace("__", "."): torch.from_numpy(v) for k, v in fx.items()
                              if k.startswith("sd__")})


def test_host_utils_match_reference_known_answers():
    from helpers import load
    from yolo_continuous_b200.utils import bbox
    fx = load("bbox_kat")
    for f in bbox.CvtFlag:
        assert np.array_equal(bbox.cvt_bbox(fx["boxes"].copy(), f), fx[f"out_{f.value}"])
    assert np.array_equal(bbox.make_grid(5, 3).numpy(), fx["grid_5_3"])
    with pytest.raises(Exception):
        bbox.cvt_bbox(fx["boxes"], type("F", (), {"value": 9})())


def test_yolo_correct_boxes_host_matches_oracle():
    from oracle import oracle as orc
    from yolo_continuous_b200 import detect
    g = np.random.default_rng(0)
    rows = np.zeros((50, 7), np.float32)
    rows[:, :2] = g.uniform(0, 0.5, (50, 2)); rows[:, 2:4] = rows[:, :2] + g.uniform(0.01, 0.5, (50, 2))
    for lb, shp in ((True, (512, 773)), (False, (480, 640)), (True, (1080, 1920))):
        want = orc.correct_boxes_rows(rows.copy(), (640, 640), shp, lb)[:, :4]
        xy, wh = (rows[:, 0:2] + rows[:, 2:4]) / 2, rows[:, 2:4] - rows[:, 0:2]
        got = detect.yolo_correct_boxes(xy, wh, (640, 640), np.array(shp), lb).astype(np.float32)
        assert np.array_equal(got, want)


def test_product_never_touches_the_oracle_or_the_reference():
    """oracle/ is test infrastructure: nothing under the package may import it (or the reference checkout), and the
    package has no CPU implementation to fall back to (CPU tensors raise, see test_product_refuses_cpu_tensors)."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for f in glob.glob(os.path.join(root, "yolo_continuous_b200", "**", "*.py"), recursive=True):
        src = open(f).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "/root/reference" in src:
            bad.append(f)
    assert not bad, bad
    for f in glob.glob(os.path.join(root, "yolo_continuous_b200", "csrc", "*.cu*")):
        assert "oracle" not in open(f).read().lower().replace("oracle computes", ""), f


def test_half_split_offsets_stay_inside_the_accumulator(tmp_path):
    """Host logic of the half-row z epilogue (csrc/yc_head_tc.cuh, half_off_for): for every head shape the offset it picks
    lets lanes 16-31 read the columns [OFF, 2 OFF) of the LAST anchor of a tile without leaving the 256-column TMEM buffer,
    covers the row (2 OFF >= no) and is one of the instantiated values; shapes it declines fall back to the whole-row
    epilogue.  Compiled with nvcc as host code: no GPU needed."""
    src = tmp_path / "off.cu"
    src.write_text(r'''
#include <cstdio>
#include "yc_head_tc.cuh"
int main() {
  int bad = 0, taken = 0;
  for (int na = 1; na <= 4; ++na)
    for (int no = 6; no <= 256; ++no) {
      if (na * no > 256) continue;              // all anchors in one tile, as launch_head_tcgen05 requires
      const int off = yc::half_off_for(no, na);
      if (!off) continue;
      ++taken;
      const bool inst = off == 4 || off == 8 || off == 16 || off == 32 || off == 43 || off == 64;
      if (!inst || 2 * off < no || off > no || (na - 1) * no + 2 * off > 256) { printf("bad: na %d no %d off %d\n", na, no, off); ++bad; }
    }
  // the shapes of the shipped heads take the half-row path
  if (yc::half_off_for(85, 3) != 43 || yc::half_off_for(6, 3) != 4 || yc::half_off_for(25, 3) != 16 || yc::half_off_for(127, 1) != 64) ++bad;
  printf("%d %d\n", bad, taken);
  return 0; }''')
    exe = tmp_path / "off"
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                           "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "yolo_continuous_b200", "csrc"),
                           str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[-2]) == 0 and int(out[-1]) > 300, out


def test_chunked_tile_order_visits_every_tile_once(tmp_path):
    """Host check of tile_coord_w (csrc/yc_head_tc.cuh): the tile order of heads with one anchor group per tile -- group major
    inside chunks of pixel tiles, the last chunk short -- is a permutation of (level, image, pixel tile, group)."""
    src = tmp_path / "order.cu"
    src.write_text(r'''
#include <cstdio>
#include <cstring>
#include <set>
#include <tuple>
#include "yc_head_tc.cuh"
int main() {
  int bad = 0;
  const int chunks[] = {1, 7, 148, 296, 100000};
  for (int bs = 1; bs <= 5; bs += 2)
    for (int ci = 0; ci < 5; ++ci) {
      yc::TcParams P;
      memset(&P, 0, sizeof(P));
      P.n_lv = 3; P.bs = bs;
      const int hw[3] = {400, 1600, 6400};
      int tiles = 0;
      for (int l = 0; l < 3; ++l) {
        P.lv[l].HW = hw[l]; P.lv[l].tiles_per_img = (hw[l] + 127) / 128; P.lv[l].n_groups = 3; P.lv[l].chunk_tiles = chunks[ci];
        P.lv[l].tile_begin = tiles; tiles += bs * P.lv[l].tiles_per_img * 3;
      }
      P.total_tiles = tiles;
      std::set<std::tuple<int, int, int, int>> seen;
      for (int t = 0; t < tiles; ++t) {
        const yc::TileCoord c = yc::tile_coord_w(P, t, 128);
        if (c.lv < 0 || c.lv > 2 || c.g < 0 || c.g > 2 || c.b < 0 || c.b >= bs || c.p0 < 0 || c.p0 >= hw[c.lv] || c.p0 % 128) ++bad;
        seen.insert(std::make_tuple(c.lv, c.b, c.p0, c.g));
      }
      if ((int)seen.size() != tiles) ++bad;
    }
  printf("%d\n", bad);
  return 0; }''')
    exe = tmp_path / "order"
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                           "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "yolo_continuous_b200", "csrc"),
                           str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)], text=True).split()[-1] == "0"
