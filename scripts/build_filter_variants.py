"""Builds A/B variants of the library that differ only in score_filter_tc.cu (a git revision of the file, or the working
copy with textual patches) into _ab/lib_<name>.so, for scripts/filter_ab.py.  Development aid."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'tensorrec_b200', 'csrc')
sys.path.insert(0, ROOT)
from tensorrec_b200.csrc import build as B  # noqa: E402


def make(name, text):
    tmp = os.path.join(CSRC, '_variant_%s.cu' % name)
    obj = '/tmp/_variant_%s.o' % name
    with open(tmp, 'w') as fh:
        fh.write(text)
    try:
        subprocess.run([B.find_nvcc()] + B.NVCC_FLAGS + ['-c', tmp, '-o', obj], check=True)
    finally:
        os.remove(tmp)
    objs = [os.path.join(CSRC, s.replace('.cu', '.o')) for s in B.SOURCES if s != 'score_filter_tc.cu'] + [obj]
    out = os.path.join(ROOT, '_ab', 'lib_%s.so' % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run([B.find_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', out] + objs, check=True)
    print(out)


def rev(r):
    return subprocess.run(['git', 'show', '%s:tensorrec_b200/csrc/score_filter_tc.cu' % r], cwd=ROOT, check=True,
                          capture_output=True, text=True).stdout


def patched(text, pairs):
    for a, b in pairs:
        assert a in text, a
        text = text.replace(a, b)
    return text


if __name__ == '__main__':
    B.build()
    head = open(os.path.join(CSRC, 'score_filter_tc.cu')).read()
    variants = {}
    for spec in sys.argv[1:]:
        name, what = spec.split('=', 1)
        variants[name] = rev(what) if what != 'WORK' else head
    for name, text in variants.items():
        make(name, text)
