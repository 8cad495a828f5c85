#!/usr/bin/env python3
"""Assemble the asset tree the reference's loaders expect -> oracle/_ref/work/ (TEST INFRASTRUCTURE ONLY).

The reference resolves every path as "../assets/..." relative to its working directory
(SURVEY.md section 8c item 12), and its checkout lacks several large blobs
(/root/reference/.MISSING_LARGE_BLOBS): the HDR skydome used by every scene, log_fence.png,
T_Trim_0x_BaseColor.png, urna.obj/jpg.  This script

  * copies the OBJ models, textures and shipped scene XMLs it needs from /root/reference/assets
    into oracle/_ref/work/assets/ (git-ignored; it travels to the GPU box so the CPU baseline can
    load the same scenes there),
  * generates deterministic stand-ins for the missing blobs (seeded; documented in DESIGN.md),
  * copies this repo's authored scene files (scenes/*.xml) next to the shipped ones.

Layout:  oracle/_ref/work/assets/...   and   oracle/_ref/work/run/   (the cwd for the reference).
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("RT_REFERENCE_DIR", "/root/reference")
WORK = os.path.join(HERE, "_ref", "work")

MODELS = ["bunny.obj", "wok.obj", "wok.mtl", "teapot.obj", "teapot.mtl", "japanese_torii_gate.obj",
          "watch-tower.obj", "log_fence.obj", "log_fence.mtl", "cube.obj", "cube.mtl"]
TEXTURES = ["Defuse_wok.png", "Stylized_Brick_basecolor.png", "Stylized_Pavement_basecolor.png",
            "Stylized_Wood_basecolor.tga", "Wood_Tower_Col.jpg"]
SCENES = ["inside_scene.xml", "different_size_scene.xml", "uniform_distributed_scene.xml"]


def write_standin_hdr(path, w=2048, h=1024):
    """Radiance .hdr, flat (non-RLE) RGBE scanlines.  stb_image decodes a file flat when the first
    scanline does not start with bytes (2, 2, <128) (lib/stb_image.h:7138-7156) and then tone-maps
    it to 8 bit in stbi_load (stb_image.h:1864-1875), exactly as it would the real 4k sky."""
    v, u = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing="ij")
    elev = 1.0 - v  # 1 = zenith row (Texture::Sample flips v, so row 0 is sampled for theta = pi)
    # sunset gradient + sun disc + banded clouds: smooth but with enough structure that a texel flip shows
    sky = np.stack([0.25 + 0.9 * elev ** 2, 0.35 + 0.5 * elev, 0.9 - 0.55 * elev], -1)
    clouds = 0.5 + 0.5 * np.sin(u * 37.0 + 3.0 * np.sin(v * 11.0)) * np.sin(v * 23.0 + 1.3)
    sky = sky * (0.65 + 0.35 * clouds[..., None])
    sun = np.exp(-(((u - 0.62) * 2.0) ** 2 + (v - 0.42) ** 2) * 900.0)
    rgb = (sky + sun[..., None] * np.array([6.0, 4.5, 2.5])).astype(np.float32)
    m = rgb.max(-1)
    e = np.ceil(np.log2(np.maximum(m, 1e-30))).astype(np.int32)
    scale = np.exp2(-e.astype(np.float64))[..., None] * 256.0
    mant = np.clip(np.floor(rgb * scale), 0, 255).astype(np.uint8)
    rgbe = np.concatenate([mant, (e + 128).astype(np.uint8)[..., None]], -1)
    assert not (rgbe[0, 0, 0] == 2 and rgbe[0, 0, 1] == 2), "first pixel would look like an RLE header"
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(rgbe.tobytes())


def write_standin_png(path, seed, size=512, base=(0.55, 0.4, 0.25)):
    from PIL import Image
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    grain = np.sin(x * 0.11 + 2.0 * np.sin(y * 0.023)) * 0.5 + 0.5
    noise = rng.random((size // 8, size // 8)).repeat(8, 0).repeat(8, 1)
    img = np.array(base)[None, None, :] * (0.6 + 0.3 * grain[..., None] + 0.1 * noise[..., None])
    Image.fromarray((np.clip(img, 0, 1) * 255).astype(np.uint8), "RGB").save(path)


def main():
    if not os.path.isdir(os.path.join(REF, "assets")):
        raise FileNotFoundError(f"{REF}/assets not present")
    assets = os.path.join(WORK, "assets")
    os.makedirs(os.path.join(assets, "textures"), exist_ok=True)
    os.makedirs(os.path.join(assets, "scenes"), exist_ok=True)
    os.makedirs(os.path.join(WORK, "run"), exist_ok=True)

    def cp(src, dst):
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)

    for fn in MODELS:
        cp(os.path.join(REF, "assets", fn), os.path.join(assets, fn))
    for fn in TEXTURES:
        cp(os.path.join(REF, "assets", "textures", fn), os.path.join(assets, "textures", fn))
    for fn in SCENES:
        cp(os.path.join(REF, "assets", "scenes", fn), os.path.join(assets, "scenes", fn))
    for fn in os.listdir(os.path.join(REPO, "scenes")):
        if fn.endswith(".xml"):
            shutil.copyfile(os.path.join(REPO, "scenes", fn), os.path.join(assets, "scenes", fn))
    hdr = os.path.join(assets, "industrial_sunset_puresky_4k.hdr")
    if not os.path.exists(hdr):
        write_standin_hdr(hdr)
    for i, fn in enumerate(["log_fence.png", "T_Trim_01_BaseColor.png", "T_Trim_02_BaseColor.png"]):
        p = os.path.join(assets, "textures", fn)
        if not os.path.exists(p):
            write_standin_png(p, seed=1234 + i, base=[(0.5, 0.36, 0.22), (0.6, 0.6, 0.62), (0.35, 0.45, 0.4)][i])
    # tiny asset set for the committed golden fixture (tests/golden): small textures keep the flattened
    # scene at a few hundred KB
    gold = os.path.join(assets, "golden")
    os.makedirs(gold, exist_ok=True)
    if not os.path.exists(os.path.join(gold, "sky_small.hdr")):
        write_standin_hdr(os.path.join(gold, "sky_small.hdr"), w=256, h=128)
    for i, fn in enumerate(["floor_small.png", "tex_small.png"]):
        if not os.path.exists(os.path.join(gold, fn)):
            write_standin_png(os.path.join(gold, fn), seed=77 + i, size=128 if i else 200,
                              base=[(0.5, 0.5, 0.55), (0.7, 0.45, 0.3)][i])
    return WORK


if __name__ == "__main__":
    print(main())
