""".fb framebuffer text I/O (src/fj_framebuffer_io.cc:46-68: "PTO" header, `resolution W H`,
`channel_count C`, one pixel per line, default ostream precision = 6 significant digits)."""
import numpy as np


def read_fb(path):
    with open(path, "r") as f:
        lines = f.read().split("\n")
    w = h = c = None
    start = None
    for i, ln in enumerate(lines):
        if ln.startswith("resolution"):
            _, w, h = ln.split()
            w, h = int(w), int(h)
        elif ln.startswith("channel_count"):
            c = int(ln.split()[1])
        elif ln.startswith("begin pixels"):
            start = i + 1
            break
    body = lines[start:start + w * h]
    arr = np.array(" ".join(body).split(), dtype=np.float32)
    return arr.reshape(h, w, c)


def write_fb(path, img):
    h, w, c = img.shape
    with open(path, "w") as f:
        f.write("#PTO Plain Text Object\n#Fujiyama Renderer FrameBuffer\n")
        f.write("resolution %d %d\nchannel_count %d\nbegin pixels\n" % (w, h, c))
        flat = img.reshape(-1, c)
        f.write("\n".join(" ".join("%g" % v for v in px) for px in flat))
        f.write("\nend pixels\n")


def rmse_per_channel(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return np.sqrt((d * d).reshape(-1, a.shape[-1]).mean(0))
