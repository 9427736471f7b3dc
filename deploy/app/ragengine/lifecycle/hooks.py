#!/usr/bin/env python3
"""PostStart / PreStop hooks. The controller hard-codes
`python3 /app/ragengine/lifecycle/hooks.py poststart|prestop` (pkg/ragengine/manifests/manifests.go:116-134).

On-disk protocol of the reference (presets/ragengine/lifecycle/manager.py:126-326), kept byte-compatible so that a PVC written
by either image restores under the other:
    <base>/systemsnapshots/<YYYY-mm-ddTHH-MM-SS>_pod-<uid8>/<index>/...      one directory per index (service /persist)
    <base>/systemsnapshots/<...>/metadata.json   {timestamp, pod_name, pod_uid, index_names, version: 1}
    <base>/LATEST -> systemsnapshots/<newest>    relative symlink; PostStart falls back to the newest directory without it
    newest 5 snapshots kept; base = $DEFAULT_VECTOR_DB_PERSIST_DIR (default /mnt/vector-db)
Exit code 0 on success / nothing to do, 1 on failure, like the reference handlers."""
import json
import os
import shutil
import sys
import time
import urllib.parse
import urllib.request
from datetime import datetime

BASE = os.environ.get("RAG_SERVICE_URL", "http://localhost:5000")
KEEP = 5


def _base_dir():
    return os.environ.get("DEFAULT_VECTOR_DB_PERSIST_DIR", "/mnt/vector-db")


def _post_ok(url, timeout=30):
    try:
        req = urllib.request.Request(url, method="POST", data=b"")
        with urllib.request.urlopen(req, timeout=timeout) as r:
            return r.status == 200
    except Exception as e:
        print(f"request failed: {url}: {e}")
        return False


def _get_indexes():
    try:
        with urllib.request.urlopen(BASE + "/indexes", timeout=5) as r:
            return json.loads(r.read() or b"[]")
    except Exception as e:
        print(f"Failed to get indexes: {e}")
        return []


def _wait_ready(seconds=60):
    for _ in range(max(1, seconds // 2)):
        try:
            with urllib.request.urlopen(BASE + "/indexes", timeout=2) as r:
                if r.status == 200:
                    return True
        except Exception:
            pass
        time.sleep(2)
    return False


def _snapshots(root):
    return sorted((d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d))), reverse=True) if os.path.isdir(root) else []


def poststart(base=None) -> int:
    base = base or _base_dir()
    link = os.path.join(base, "LATEST")
    if not _wait_ready():
        print("ERROR: Service did not become ready")
        return 1
    if os.path.islink(link) and os.path.exists(link):
        latest = os.path.realpath(link)
    else:
        root = os.path.join(base, "systemsnapshots")
        snaps = _snapshots(root)
        if not snaps:
            print("No previous snapshots found")
            return 0
        latest = os.path.join(root, snaps[0])
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(os.path.join("systemsnapshots", snaps[0]), link)        # recreate LATEST, relative to base
    meta = os.path.join(latest, "metadata.json")
    if not os.path.exists(meta):
        print("No metadata found in snapshot")
        return 0
    try:
        names = json.load(open(meta)).get("index_names", [])
        loaded = 0
        for name in names:
            q = urllib.parse.urlencode({"path": os.path.join(latest, name), "overwrite": "true"})
            if _post_ok(f"{BASE}/load/{urllib.parse.quote(name, safe='')}?{q}"):
                loaded += 1
                print(f"Loaded: {name}")
            else:
                print(f"Failed to load: {name}")
        print(f"PostStart complete: {loaded}/{len(names)} indexes loaded")
        return 0
    except Exception as e:
        print(f"ERROR in PostStart handler: {e}")
        return 1


def prestop(base=None, keep=KEEP) -> int:
    base = base or _base_dir()
    root = os.path.join(base, "systemsnapshots")
    os.makedirs(root, exist_ok=True)
    uid = (os.environ.get("POD_UID", "unknown") or "unknown")[:8]
    snap_name = f"{datetime.now().strftime('%Y-%m-%dT%H-%M-%S')}_pod-{uid}"
    snap = os.path.join(root, snap_name)
    os.makedirs(snap, exist_ok=True)
    try:
        names = _get_indexes()
        if not names:
            print("No indexes to persist")
            shutil.rmtree(snap)
            return 0
        saved = [n for n in names
                 if _post_ok(f"{BASE}/persist/{urllib.parse.quote(n, safe='')}?{urllib.parse.urlencode({'path': os.path.join(snap, n)})}")]
        if not saved:
            print("No indexes were successfully saved")
            shutil.rmtree(snap)
            return 1
        meta = {"timestamp": datetime.now().isoformat(), "pod_name": os.environ.get("POD_NAME", "unknown"), "pod_uid": uid,
                "index_names": saved, "version": 1}
        with open(os.path.join(snap, "metadata.json"), "w") as f:
            json.dump(meta, f, indent=2)
        link = os.path.join(base, "LATEST")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(os.path.join("systemsnapshots", snap_name), link)
        for old in _snapshots(root)[keep:]:
            shutil.rmtree(os.path.join(root, old))
            print(f"Deleted old snapshot: {old}")
        print(f"PreStop complete: {len(saved)} indexes saved to {snap}")
        return 0
    except Exception as e:
        print(f"ERROR in PreStop handler: {e}")
        shutil.rmtree(snap, ignore_errors=True)
        return 1


def main():
    if len(sys.argv) != 2 or sys.argv[1].lower() not in ("poststart", "prestop"):
        print("Usage: hooks.py [poststart|prestop]")
        sys.exit(1)
    sys.exit(poststart() if sys.argv[1].lower() == "poststart" else prestop())


if __name__ == "__main__":
    main()
