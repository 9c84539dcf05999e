#!/bin/bash
# The reference's pipeline (pandelos.sh:46-79: calculate_k.py -> java Pangenes -> netclu_ng.py -> grep/sed/sort) with the
# three programs replaced by this package's native ones; same arguments, same <out_prefix>.clus.
#
#   pandelos_b200/pandelos.sh dataset.faa out_prefix
#
# The similarity stage runs on the GPUs of this host (all of them; PD_DEVICES=n caps it), the other two on the CPU.
# PD_PANGENES="command" puts another program with the reference's -i/-k/-o flags in its place (e.g. the reference's own
# `java ... infoasys.cli.pangenes.Pangenes` line over the JNI shim).  Families are sorted in byte order
# (LC_ALL=C), where the reference's script leaves the order to the caller's locale.
set -e
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
idb="$1"
oprefix="$2"
if [ ! -f "$idb" ] || [ -z "$oprefix" ]; then
	echo "usage is: pandelos.sh dataset.faa out_prefix" >&2
	exit 1
fi
pangenes="${PD_PANGENES:-$here/pangenes}"
tools="calculate_k netclu_cc"
[ -z "$PD_PANGENES" ] && tools="$tools pangenes"
for tool in $tools; do
	if [ ! -x "$here/$tool" ]; then
		echo "ERROR: $here/$tool not built (python -m pandelos_b200.build)" >&2
		exit 1
	fi
done
tmp=`mktemp -d -p ./ $(basename "$idb" .faa).XXXXXX`
trap 'rm -rf "$tmp"' EXIT

"$here/calculate_k" "$idb" > "$tmp/k.txt"
k=`grep -E "^k =" "$tmp/k.txt" | sed s/k\ =\ //g`
k=$((k))
echo "k = $k"
$pangenes -i "$idb" -k $k -o "$tmp/net" > "$tmp/pangenes.txt"
grep "Total cost" "$tmp/pangenes.txt" || true
"$here/netclu_cc" "$idb" "$tmp/net" -g > "$tmp/fam.txt"
grep "F{ " "$tmp/fam.txt" | sed s/F{\ //g | sed s/}//g | sed s/\ \;//g | LC_ALL=C sort | uniq > "${oprefix}.clus"
echo "gene families written to ${oprefix}.clus"
