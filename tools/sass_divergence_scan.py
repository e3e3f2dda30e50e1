#!/usr/bin/env python
"""Static scan of the built kernels for warp collectives that part of a warp can skip.

ptxas guards every shuffle / vote that may be reached by a diverged warp with
`BRA.DIV` (branch to a WARPSYNC.COLLECTIVE slow path that waits for ALL lanes of
the mask).  A lane-predicated branch (`@P BRA`, not `BRA.U`) a few instructions
before such a guard means some lanes jump over the collective while the others
wait for them: a deadlock on the device.  That is the signature the round-1 hang
of the interior-point kernel left in the SASS (a shuffle reduction inside a
short-circuited `&&`); the host emulator (tests/emu) finds the same bug class
dynamically.  Loop back-edges right before a collective show up as (harmless)
hits too: read the listing, do not gate on it.

    python tools/sass_divergence_scan.py [object files ...]   (default: .scratch/obj/*.o)
"""

import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, ".scratch", "obj", "*.o")))
INS = re.compile(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);")
total = 0
for obj in objs:
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    func, ins = None, []

    def flush():
        global total
        hits = []
        for i, (addr, text) in enumerate(ins):
            if not text.startswith("BRA.DIV"):
                continue
            for j in range(max(0, i - 4), i):
                if re.match(r"@!?P\d BRA", ins[j][1]):
                    back = int(re.search(r"0x([0-9a-f]+)", ins[j][1]).group(1), 16) <= ins[j][0]
                    hits.append(f"    {ins[j][0]:#x} {ins[j][1]:<28} before {addr:#x} {text}"
                                + ("   (loop back-edge)" if back else "   <-- lanes can skip the collective"))
        if func:
            print(f"{os.path.basename(obj)}: {func}: {len(ins)} instructions, {len(hits)} hit(s)")
            for h in hits:
                print(h)
            total += sum("skip" in h for h in hits)

    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            flush()
            func, ins = m.group(1), []
            continue
        m = INS.match(line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    flush()
print(f"{total} collective(s) that part of a warp can skip")
