"""Stage the UNMODIFIED reference modules for the CPU reference arm - TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/stage_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

Copies the four files of the reference's hot path (models/{__init__,model,layer,sublayer,allennlp_beamsearch}.py) from the
reference checkout into oracle/_ref/models/, byte for byte (a sha256 manifest is written next to them).  oracle/_ref/ is
git-ignored - reference sources never enter this repository's history - but it is NOT gpurun-ignored, so the copy travels
to the GPU box, where /root/reference does not exist, and `bench.py --impl reference` can time the reference's own code on
the box's host cores (kind "reference"); without the copy that arm falls back to the oracle port (kind "port").
The product never imports anything from here.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref', 'models')
FILES = ['__init__.py', 'model.py', 'layer.py', 'sublayer.py', 'allennlp_beamsearch.py']


def stage(ref=None):
    ref = ref or os.environ.get('DLSG_REFERENCE', '/root/reference')
    src = os.path.join(ref, 'models')
    if not os.path.isdir(src):
        return None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(os.path.join(DST, f), 'rb').read()).hexdigest()
    with open(os.path.join(HERE, '_ref', 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': src, 'sha256': manifest}, fh, indent=1)
    return DST


if __name__ == '__main__':
    d = stage(sys.argv[1] if len(sys.argv) > 1 else None)
    print(d if d else 'reference checkout not found: nothing staged')
