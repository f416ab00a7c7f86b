"""Generates tests/golden/ref_laplace2d_stdout.txt -- the standard output of the reference's own C program
MPI/Laplace/c/laplace2d.c (4320 x 4320, 1000 Jacobi iterations, the residual printed every 100), compiled unmodified by
`make -C oracle ref` into oracle/_ref/laplace2d.  About 80 s on one core.  Run in the authoring container."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

if __name__ == "__main__":
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "laplace2d")], capture_output=True, text=True, check=True).stdout
    path = os.path.join(HERE, "ref_laplace2d_stdout.txt")
    open(path, "w").write(out)
    print("wrote", path, len(out.splitlines()), "lines")
