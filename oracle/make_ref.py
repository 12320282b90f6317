"""Recipe for oracle/_ref/: the reference's OWN, UNMODIFIED source files of the hot path, taken from where they lie under
/root/reference into oracle/_ref/src/ (git-ignored, so the repo's history never holds reference code; not gpurun-ignored, so the
directory travels to the GPU box like the built .so files).  TEST / BASELINE INFRASTRUCTURE ONLY: `bench.py --impl reference` and the
`cpu_baseline` leg time these files through oracle/ref_shim.py; nothing under opentf_b200/ imports them.

    python -m oracle.make_ref        (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil

REF = '/root/reference/src'
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref', 'src')
FILES = ['mdl/__init__.py', 'mdl/ntf.py', 'mdl/fnn.py', 'mdl/bnn.py', 'mdl/earlystopping.py', 'mdl/tntf.py']  # the files on the path


def make():
    """-> number of files placed (0 when the reference tree is not mounted: the GPU box uses what was placed here)."""
    if not os.path.isdir(REF): return 0
    n = 0
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        if not os.path.exists(src): continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    return n


if __name__ == '__main__':
    print(f'{make()} reference files -> {DST}')
