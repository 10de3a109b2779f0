"""SASS bytes per device function of one kernel object (instruction-cache footprint study): python tools/code_size.py inst_15_3.o"""
import subprocess, sys, re
obj = sys.argv[1]
out = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
rows = []; on = False
for line in out.splitlines():
    if line.startswith(".section .symtab"): on = True; continue
    if on and line.startswith(".section"): break
    a = line.split()
    if on and len(a) >= 7 and a[3] in ("0x2", "0x12", "0x22") and "$" in a[-1]:
        rows.append((int(a[2], 16), int(a[1], 16), a[-1].split("$")[-1]))
    elif on and len(a) >= 7 and a[3] == "0x12":
        rows.append((int(a[2], 16), int(a[1], 16), a[-1]))
rows.sort(reverse=True); tot = 0
for sz, off, n in rows:
    name = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"nmpc::LayT<(\d+), (\d+), (\w+)>", r"L", name); name = re.sub(r"\(.*", "", name)
    print(f"{sz / 1024:8.1f} KB  {name[:90]}"); tot += sz
print(f"{tot / 1024:8.1f} KB total")
