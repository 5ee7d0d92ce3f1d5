"""Throughput of the batched dataset driver on BASELINE configs[0] (src/cli.py --n 1000 --seed 0).
  emit   (needs /root/reference, CPU only): the reference's rule logic with recording seams, samples/s on the host cores
  own    (needs a GPU; uses the recorded fixture tests/golden/config0.json.gz): render + overview + device PNG + files on disk
usage: python tools/driver_probe.py emit|own [out_dir]"""
import gzip
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
mode = sys.argv[1] if len(sys.argv) > 1 else 'emit'
out = sys.argv[2] if len(sys.argv) > 2 else '/tmp/rig_driver_probe'
shutil.rmtree(out, ignore_errors=True)
os.makedirs(out, exist_ok=True)
from reasoning_image_generation_b200 import driver

if mode == 'emit':
    workers = os.cpu_count() or 1
    t = time.perf_counter()
    n = ok = panels = 0
    for em in driver.emit_dataset('/root/reference/src', out, 1000, 0, 3, workers=workers):
        n += 1
        ok += bool(em.ok)
        panels += em.scene.n_panels if em.ok else 0
    dt = time.perf_counter() - t
    print(json.dumps({'mode': 'emit', 'samples': n, 'ok': ok, 'panels': panels, 'seconds': dt, 'samples_per_s': n / dt, 'workers': workers}))
else:
    from test_config0 import _emissions
    cfg = json.loads(gzip.open(os.path.join(ROOT, 'tests', 'golden', 'config0.json.gz')).read())
    for write in (False, True):
        ems = _emissions(cfg, out)
        w = driver.DatasetWriter(512, 512, device=0, batch_panels=4096, write_files=write)
        for em in ems[:64]:
            w.add(em)
        w.flush()                                  # warm-up: contexts, scratch, page-locked buffers
        t = time.perf_counter()
        for em in ems:
            w.add(em)
        w.close()
        dt = time.perf_counter() - t
        print(json.dumps({'mode': 'own', 'write_files': write, 'samples': len(ems), 'panels': sum(e.scene.n_panels for e in ems), 'seconds': dt,
                          'samples_per_s': len(ems) / dt}))
