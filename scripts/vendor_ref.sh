#!/bin/bash
# Vendor the UNMODIFIED reference Python sources the baseline arms and the install() test import into the git-ignored
# baseline/_ref/ (it travels to the GPU box with the snapshot like the built .so; /root/reference does not exist there).
# Nothing under magnet_b200/ may import from it — it is the thing measured against, never the product.
set -eu
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
if [ ! -d "$SRC/models" ]; then echo "vendor_ref: $SRC not present — keeping $DST as is"; exit 0; fi
rm -rf "$DST"; mkdir -p "$DST"
for d in models utils data; do
  mkdir -p "$DST/$d"
  (cd "$SRC/$d" && find . -name '*.py' -o -name '*.json' | while read -r f; do mkdir -p "$DST/$d/$(dirname "$f")"; cp "$f" "$DST/$d/$f"; done)
done
(cd "$SRC" && find models utils data -name '*.py' -o -name '*.json' | sort | xargs sha256sum) > "$DST/SHA256SUMS"
echo "vendored $(wc -l < "$DST/SHA256SUMS") files from $SRC into $DST"
