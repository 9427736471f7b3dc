#!/usr/bin/env python3
"""PostStart / PreStop hooks. The controller hard-codes
`python3 /app/ragengine/lifecycle/hooks.py poststart|prestop` (pkg/ragengine/manifests/manifests.go:116-134);
this keeps the reference's snapshot protocol (presets/ragengine/lifecycle/manager.py:126-326):
  <persist_dir>/systemsnapshots/<ts>_pod-<uid8>/<index>/ , metadata.json {index_names, version: 1},
  LATEST symlink, newest 5 snapshots kept -- over the service's own /indexes, /persist, /load routes."""
import json
import os
import shutil
import sys
import time
import urllib.parse
import urllib.request

BASE = os.environ.get("RAG_SERVICE_URL", "http://localhost:5000")
ROOT = os.path.join(os.environ.get("DEFAULT_VECTOR_DB_PERSIST_DIR", "storage"), "systemsnapshots")
KEEP = 5


def _call(method, path, timeout=300):
    req = urllib.request.Request(BASE + path, method=method, data=b"" if method == "POST" else None)
    with urllib.request.urlopen(req, timeout=timeout) as r:
        return json.loads(r.read() or b"null")


def _wait_ready(seconds=600):
    t0 = time.time()
    while time.time() - t0 < seconds:
        try:
            return _call("GET", "/indexes", timeout=5)
        except Exception:
            time.sleep(2)
    raise SystemExit("service did not become ready")


def poststart():
    _wait_ready()
    latest = os.path.join(ROOT, "LATEST")
    meta = os.path.join(latest, "metadata.json")
    if not os.path.exists(meta):
        print("no snapshot to restore")
        return
    for name in json.load(open(meta)).get("index_names", []):
        q = urllib.parse.urlencode({"path": os.path.join(os.path.realpath(latest), name), "overwrite": "true"})
        _call("POST", f"/load/{urllib.parse.quote(name, safe='')}?{q}")
        print(f"restored index {name}")


def prestop():
    names = _call("GET", "/indexes")
    if not names:
        return
    uid = (os.environ.get("POD_UID", "unknown") or "unknown")[:8]
    snap = os.path.join(ROOT, f"{time.strftime('%Y%m%d-%H%M%S')}_pod-{uid}")
    os.makedirs(snap, exist_ok=True)
    for name in names:
        q = urllib.parse.urlencode({"path": os.path.join(snap, name)})
        _call("POST", f"/persist/{urllib.parse.quote(name, safe='')}?{q}")
    json.dump({"index_names": names, "version": 1}, open(os.path.join(snap, "metadata.json"), "w"))
    tmp = os.path.join(ROOT, "LATEST.tmp")
    if os.path.lexists(tmp):
        os.remove(tmp)
    os.symlink(os.path.basename(snap), tmp)
    os.replace(tmp, os.path.join(ROOT, "LATEST"))
    snaps = sorted(d for d in os.listdir(ROOT) if os.path.isdir(os.path.join(ROOT, d)) and not os.path.islink(os.path.join(ROOT, d)))
    for old in snaps[:-KEEP]:
        shutil.rmtree(os.path.join(ROOT, old), ignore_errors=True)


if __name__ == "__main__":
    {"poststart": poststart, "prestop": prestop}[sys.argv[1]]()
