"""Counts of the SASS mnemonics that identify the Blackwell paths (tcgen05 MMA, TMEM load/store, TMA bulk-tensor loads, warp MMA)
per kernel of the built objects:  python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections, glob, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "LDGSTS", "HMMA", "SYNCS", "MUFU.EX2", "FFMA2", "RED.E.ADD", "ATOMG"]
print("kernel".ljust(64) + " ".join(k.rjust(9) for k in keys))
for obj in sorted(glob.glob(os.path.join(root, "clear_vae_b200", "_build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, cnt = None, None
    rows = []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, cnt))
            name, cnt = m.group(1), collections.Counter()
            continue
        if name:
            for k in keys:
                if re.search(r"\b" + re.escape(k), line):
                    cnt[k] += 1
    if name:
        rows.append((name, cnt))
    print(f"# {os.path.basename(obj)}")
    for name, cnt in rows:
        if not any(cnt[k] for k in keys[:8]):
            continue
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        dem = re.sub(r"\(anonymous namespace\)::|<unnamed>::|\((int|bool)\)", "", dem).split("(")[0].replace("void ", "")
        print(dem[:63].ljust(64) + " ".join(str(cnt[k]).rjust(9) for k in keys))
