"""The single-lattice (AA) kernels read the lattice through a const __restrict__ alias with non-coherent loads while storing
through another pointer to the same array (lbm_aa_kernels.inl, "Aliasing note").  That is only sound as long as the compiler
keeps every lattice load of a thread ahead of its lattice stores.  This test reads the SASS of the built kernels (no GPU
needed) and fails if, on any straight-line path, a global load follows a global store -- the situation a future toolchain
could create under the restrict promise, and which the in-place scheme has no second copy to reveal."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = [os.path.join(ROOT, "mglc_b200", "csrc", "build", n) for n in ("lbm_aa_fast.o", "lbm_aa.o")]


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name and "/*" in line:
            body.append(line)
    if name:
        yield name, body


@pytest.mark.parametrize("obj", OBJS)
def test_every_lattice_load_precedes_the_stores_on_each_path(obj):
    if not os.path.exists(obj):
        pytest.skip("library not built")
    seen = 0
    for name, body in kernels(obj):
        if "k_aa_even" not in name and "k_aa_odd" not in name:
            continue
        seen += 1
        # address -> instruction text
        ins = []
        for line in body:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        # walk every path from the entry: state = (index, stored?)
        work, visited = [(0, False)], set()
        while work:
            i, stored = work.pop()
            while i < len(ins):
                if (i, stored) in visited:
                    break
                visited.add((i, stored))
                text = ins[i][1]
                pred = text.startswith("@")
                op = text.split()[1] if pred else text.split()[0]
                if op.startswith("STG"):
                    stored = True
                elif op.startswith("LDG") and stored and "rho_lid" not in text:
                    # the lid plane (a separate array) is read by the first thread instructions only; any LDG here is the lattice
                    raise AssertionError(f"{name}: global load after a global store on one path: {ins[i]}")
                if op in ("EXIT", "RET") and not pred:
                    break
                if op == "BRA":
                    tgt = re.search(r"0x([0-9a-f]+)", text)
                    if tgt and int(tgt.group(1), 16) in addr_index:
                        j = addr_index[int(tgt.group(1), 16)]
                        if j != i:
                            work.append((j, stored))
                    if not pred:
                        break
                i += 1
    assert seen >= 2
