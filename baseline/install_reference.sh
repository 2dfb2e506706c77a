#!/bin/bash
# Installs the UNMODIFIED reference package (Teichlab/Genes2Genes, /root/reference) into baseline/_ref so that
# `bench.py --impl reference` and the cpu_baseline leg can time the reference's own CPU path on the GPU box
# (baseline/_ref is git-ignored but travels with gpurun).  The reference builds with flit_core, which this
# offline image does not have: a scratch copy under /tmp gets a setuptools [build-system] table instead; the
# package sources (genes2genes/*.py) are installed byte for byte.  No dependencies are installed: the
# plotting / enrichment imports are stubbed at run time (oracle/ref_runner.py).
set -euo pipefail
SRC=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
TMP=$(mktemp -d /tmp/g2g_ref_XXXX)
cp -r "$SRC"/. "$TMP"/
python - "$TMP/pyproject.toml" <<'PY'
import re, sys
p = sys.argv[1]
s = open(p).read()
s = re.sub(r'\[build-system\].*?\n\n', '[build-system]\nrequires = ["setuptools"]\nbuild-backend = "setuptools.build_meta"\n\n', s, flags=re.S)
s = s.replace('dynamic = ["version", "description"]', 'version = "0.2.0"\ndescription = "Genes2Genes (reference, unmodified sources)"')
s += '\n[tool.setuptools]\npackages = ["genes2genes"]\n'
open(p, 'w').write(s)
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --target "$HERE/_ref" "$TMP" 2>&1 | tail -2
rm -rf "$TMP"
# byte-for-byte check of the installed sources against the reference tree
for f in "$SRC"/genes2genes/*.py; do cmp -s "$f" "$HERE/_ref/genes2genes/$(basename "$f")" || { echo "MISMATCH $f"; exit 1; }; done
echo "installed $(ls "$HERE/_ref/genes2genes"/*.py | wc -l) unmodified source files into $HERE/_ref"
