"""Probe: does nvJPEG batched decode scale across host threads (one context each)?"""
import os, sys, threading, time
import numpy as np, cv2
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import sift_features_b200 as sf
w, h = 1920, 1080
bird = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "bird_gray.npy"))
tile = np.tile(bird, (h // bird.shape[0] + 1, w // bird.shape[1] + 1))[:h, :w]
jpegs = [cv2.imencode(".jpg", np.roll(tile, 37 * i, 1), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for i in range(128)]
for T in (1, 2, 4):
    exs = [sf.Extractor(w, h, 32) for _ in range(T)]
    for ex in exs: ex.sift_jpeg(jpegs)
    def work(ex):
        for _ in range(3): ex.sift_jpeg(jpegs)
    ts = [threading.Thread(target=work, args=(ex,)) for ex in exs]
    t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]; dt = time.perf_counter() - t0
    print(T, "threads:", 3 * 128 * T / dt, "images/s", flush=True)
    for ex in exs: ex.close()
