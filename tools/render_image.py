"""Render a scene to an image file the way the reference displays it (accumulator * 1/(spp+passes) through
RGBF32_to_RGB8, renderer.cpp:119,127-129) — the headless counterpart of the reference's window.
usage: render_image.py scene.rtscene[.gz] out.(ppm|png) [pt|whitted] [W H] [spp] [camx camy camz tx ty tz]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api


def write_image(path, rgb8):
    """rgb8: (H, W) uint32 0x00RRGGBB as Surface::pixels holds them"""
    h, w = rgb8.shape
    img = np.stack([(rgb8 >> 16) & 255, (rgb8 >> 8) & 255, rgb8 & 255], -1).astype(np.uint8)
    if path.lower().endswith(".png"):
        from PIL import Image
        Image.fromarray(img, "RGB").save(path)
    else:
        with open(path, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (w, h))
            f.write(img.tobytes())


def main(argv):
    scene, out = argv[0], argv[1]
    integ = abi.RT_INTEGRATOR_WHITTED if len(argv) > 2 and argv[2] == "whitted" else abi.RT_INTEGRATOR_PATH
    W, H = (int(argv[3]), int(argv[4])) if len(argv) > 4 else (1280, 720)
    spp = int(argv[5]) if len(argv) > 5 else 64
    if not os.path.exists(scene):
        scene = os.path.join(ROOT, "oracle", "_ref", "scenes", scene + ".rtscene.gz")
    sc = api.open_scene(rtb.FlatScene.load(scene))
    r = api.GpuRenderer(sc, integ, W, H).Init()
    if len(argv) > 11:
        r.camera.SetCameraState([float(x) for x in argv[6:9]], [float(x) for x in argv[9:12]])
    if integ == abi.RT_INTEGRATOR_PATH:
        r.render(spp)
        px = r.screen_pixels(scale=1.0 / (spp + 1))  # spp counter after `spp` Ticks = spp + 1 - passes; display scale 1/(spp+passes)
    else:
        r.Tick(0)
        px = r.screen_pixels()
    write_image(out, px)
    print(out, px.shape, "mean 8-bit level", float(np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255]).mean()))


if __name__ == "__main__":
    main(sys.argv[1:])
