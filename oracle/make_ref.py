"""TEST / BASELINE INFRASTRUCTURE -- recipe that vendors the UNMODIFIED Python reference into oracle/_ref/ (git-ignored, so it
never enters the history; not gpurun-ignored, so it travels to the GPU box with the snapshot like the built .so files).

    python oracle/make_ref.py          # run by __graft_entry__.build() whenever /root/reference is present

Copies only the importable Python modules of the hot path (models/, util/, options/ -- no CUDA sources, images, scripts or
checkpoints).  Used (i) by bench.py `--impl reference` / `cpu_baseline` to time the reference's own Vid2VidModelG.inference
on the host cores (kind "reference"), and (ii) by the GPU-box test that swaps vid2vid_b200.networks into the reference's
Vid2VidModelG.  oracle/ref_shim.py finds the tree here when /root/reference does not exist."""
import os
import shutil
import sys

SRC = os.environ.get('V2V_REFERENCE_SRC', '/root/reference')
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def make():
    if not os.path.isdir(os.path.join(SRC, 'models')):
        return None
    n = 0
    for top in ('models', 'util', 'options'):
        for root, dirs, files in os.walk(os.path.join(SRC, top)):
            rel = os.path.relpath(root, SRC)
            for f in files:
                if f.endswith('.py'):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
                    n += 1
    open(os.path.join(DST, 'SOURCE.txt'), 'w').write('copied unmodified from %s by oracle/make_ref.py (%d files)\n' % (SRC, n))
    return DST


if __name__ == '__main__':
    print(make() or 'reference tree not found at %s' % SRC, file=sys.stderr if not os.path.isdir(SRC) else sys.stdout)
