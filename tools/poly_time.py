import time, numpy as np, sys
sys.path.insert(0, '.')
from motes_b200 import pipeline, synth
for cfg, F in (("C1", 1), ("C1", 16), ("C2", 1), ("C2", 32)):
    frame, axes, header = synth.make_frame(synth.config_spec(cfg, seed=3))
    args = synth.default_args(sky_order=2)
    data = np.repeat(frame["data"][None], F, 0); errs = np.repeat(frame["errs"][None], F, 0); qual = np.repeat(frame["qual"][None], F, 0)
    for rep in range(2):
        t = time.perf_counter()
        out = pipeline.run_frames(data.copy(), errs.copy(), qual, args, header["seeing"], header["pixel_resolution"], seeds=list(range(F)), axes_dict=axes,
                                  outputs=("spectra", "frame_status"))
        dt = time.perf_counter() - t
    print(cfg, F, "order 2:", round(dt * 1e3, 1), "ms wall incl. host copies", out["frame_status"][:2], flush=True)
