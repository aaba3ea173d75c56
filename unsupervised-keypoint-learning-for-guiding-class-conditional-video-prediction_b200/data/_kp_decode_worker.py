"""Decode worker process of the input pipeline (started by data/base_dataloader.py as a plain subprocess).

Stand-alone on purpose: imports only numpy and PIL, so that a worker starts in a fraction of a second and never touches
CUDA.  Protocol on stdin / stdout: 4-byte little-endian length + pickle.  A request is a list of
(staging file, byte offset, jpeg path, w, h); the worker decodes each JPEG with PIL - the reference's decoder
(data/image_pair_dataloader.py:91-92) - and writes the RGB bytes into the shared staging file at the offset; the reply is
None or the repr of the first error.
"""
import mmap
import pickle
import struct
import sys

import numpy as np
from PIL import Image


def main():
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    maps = {}
    while True:
        hdr = inp.read(4)
        if len(hdr) < 4:
            return
        tasks = pickle.loads(inp.read(struct.unpack("<I", hdr)[0]))
        err = None
        for stage, off, path, w, h in tasks:
            try:
                view = maps.get(stage)
                if view is None:
                    with open(stage, "r+b") as fh:
                        view = maps[stage] = np.frombuffer(mmap.mmap(fh.fileno(), 0), np.uint8)
                    for old in [k for k in maps if k != stage and k.rsplit("_", 1)[0] == stage.rsplit("_", 1)[0]]:
                        del maps[old]              # the slot was re-created larger
                with Image.open(path) as im:
                    px = np.asarray(im if im.mode == "RGB" else im.convert("RGB"))
                if px.shape != (h, w, 3):
                    raise ValueError("%s: decoded %s, header said %s" % (path, px.shape, (h, w, 3)))
                view[off:off + px.size] = px.reshape(-1)
            except Exception as exc:        # reported to the parent, which raises
                err = err or repr(exc)
        reply = pickle.dumps(err)
        out.write(struct.pack("<I", len(reply)) + reply)
        out.flush()


if __name__ == "__main__":
    main()
