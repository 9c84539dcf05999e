"""Pins netclu_cc -g on a network whose Girvan-Newman split the reference's netclu_ng.py cannot finish in one run
(mycoplasma64: one component of 11,325 genes; a networkx betweenness pass on it takes ~35 minutes, the whole split
removes 1,498 edges one pass each), by checking the native split LEVEL BY LEVEL against networkx:

netclu_ng.py:101-117 is a recursion whose every level depends only on (the root graph, the node list it is handed):
`snet.subgraph(com)` of a view is a view of the ROOT graph filtered by set(com) (networkx Graph.subgraph), so
`next(girvan_newman(pnet.subgraph(com)))` is exactly what the script computes when it reaches `com` — for the connected
component itself (the set connected_components yields) at the top, for the sorted member list below.  The native trace
(PD_NETCLU_TRACE=2: `rm a b` per removed edge, the script's `gn ([..], [..])` per split) names the input of every level
(the union of the two halves) and the output; if networkx maps every input to the same output, the two recursions are the
same tree.  Levels are independent, so they are checked in parallel, the small ones in seconds.

    python tests/golden/verify_netclu_splits.py in.net trace.err out.jsonl [--workers 6] [--max-nodes 7000]
        one JSON line per level: {"level", "nodes", "removed", "ok", "seconds"}; levels above --max-nodes are skipped
    python tests/golden/verify_netclu_splits.py in.net trace.err out.jsonl --sequence LEVEL [--workers 6] [--from-step J]
        one level edge by edge (for a level too large to finish in one go): a line per removed edge compared with the
        trace.  Steps are independent too — step j starts from girvan_newman's working copy with the trace's first j edges
        removed in order (what the copy is if the steps before agreed) — so they run in parallel; after the last one the
        components of what is left are compared with the trace's two halves.
    python tests/golden/verify_netclu_splits.py in.net trace.err out.clus --clus in.faa
        the `.clus` (pandelos.sh:79) assembled WITHOUT the native tool's own decisions: connected components by networkx,
        the collision test of netclu_ng.py:79-96 restated here, the trace only as the list of splits — and checked for
        consistency on the way (a component or half with a collision must be the input of a level, one without must not)
    python tests/golden/verify_netclu_splits.py in.net --first-edge
        the first edge networkx removes from the largest component
"""
import ast
import json
import sys
import time

import networkx as nx
from networkx.algorithms.community.centrality import girvan_newman


def read_pnet(path):
    """The graph as netclu_ng.py:43-56 builds it (node and adjacency insertion order included)."""
    inodes = set()
    pnet = nx.Graph()
    for line in open(path):
        cols = line.strip().split("\t")
        a, b, w = int(cols[0]), int(cols[1]), float(cols[2])
        if a not in inodes:
            inodes.add(a)
            pnet.add_node(a)
        if b not in inodes and a != b:
            inodes.add(b)
            pnet.add_node(b)
        if a != b:
            pnet.add_edge(a, b, weight=w)
            pnet.add_edge(b, a, weight=w)
    return pnet


def read_trace(path):
    """[(removed edges [(a, b)], first half, second half)] per `gn` line, in trace order."""
    levels, rm = [], []
    for line in open(path):
        if line.startswith("rm "):
            _, a, b = line.split()
            rm.append((int(a), int(b)))
        elif line.startswith("gn ("):
            first, second = ast.literal_eval(line[3:])
            levels.append((rm, first, second))
            rm = []
    return levels


PNET = None
TOP = None
LEVELS = None


def level_input(i):
    _, first, second = LEVELS[i]
    members = sorted(first + second)
    return TOP.get(tuple(members), members)   # the component's own set at the top of a recursion, the sorted list below


def check_level(i):
    t = time.time()
    rm, first, second = LEVELS[i]
    got = tuple(sorted(c) for c in next(girvan_newman(PNET.subgraph(level_input(i)))))
    return {"level": i, "nodes": len(first) + len(second), "removed": len(rm), "ok": got == (first, second), "seconds": round(time.time() - t, 1)}


SEQ_LEVEL = None


def check_step(step):
    t = time.time()
    rm, first, second = LEVELS[SEQ_LEVEL]
    g = PNET.subgraph(level_input(SEQ_LEVEL)).copy().to_undirected()   # girvan_newman's own working copy
    for e in rm[:step]:
        g.remove_edge(*e)
    if step == len(rm):
        got = tuple(sorted(c) for c in nx.connected_components(g))
        return {"level": SEQ_LEVEL, "step": "halves", "ok": got == (first, second), "seconds": round(time.time() - t, 1)}
    bet = nx.edge_betweenness_centrality(g)
    e = max(bet, key=bet.get)
    return {"level": SEQ_LEVEL, "step": step, "of": len(rm), "removed": list(e), "trace": list(rm[step]), "ok": tuple(e) == rm[step],
            "components_before": nx.number_connected_components(g), "seconds": round(time.time() - t, 1)}


def assemble_clus(faa, out_path):
    names, genome = [], []
    for i, line in enumerate(open(faa)):          # netclu_ng.py:17-30: every even line is a header
        if i % 2 == 0:
            cols = line.strip().split("\t")
            genome.append(cols[0])
            names.append(cols[1])

    def collision(members):                       # get_max_collision(...) > 0
        by_genome = {}
        for s in members:
            by_genome.setdefault(genome[s], []).append(s)
        return any(not PNET.has_edge(a, b) for v in by_genome.values() for a in v for b in v if a != b)

    inputs = {tuple(sorted(first + second)) for _, first, second in LEVELS}
    families = []

    def settle(members, what):
        key = tuple(sorted(members))
        if collision(key):
            assert key in inputs, "%s with a collision is not split in the trace: %r..." % (what, key[:5])
        else:
            assert key not in inputs, "%s without a collision is split in the trace: %r..." % (what, key[:5])
            families.append(key)

    in_net = set()
    for comp in nx.connected_components(PNET):
        in_net |= comp
        settle(comp, "component")
    for _, first, second in LEVELS:
        settle(first, "half")
        settle(second, "half")
    lines = ["F{ " + " ; ".join(names[s] for s in fam) + "}" for fam in families]
    lines += ["F{ " + names[s] + " }" for s in range(len(names)) if s not in in_net]
    clus = sorted(set(ln.replace("F{ ", "").replace("}", "").replace(" ;", "") for ln in lines))
    with open(out_path, "w") as f:
        f.write("".join(c + "\n" for c in clus))
    print("families", len(families), "outside the network", len(names) - len(in_net), "clus lines", len(clus))


def main():
    global PNET, TOP, LEVELS
    PNET = read_pnet(sys.argv[1])
    if "--first-edge" in sys.argv:
        g = PNET.subgraph(max(nx.connected_components(PNET), key=len)).copy().to_undirected()
        bet = nx.edge_betweenness_centrality(g)
        e = max(bet, key=bet.get)
        print("first", e[0], e[1], repr(bet[e]), "nodes", g.number_of_nodes(), "edges", g.number_of_edges())
        return
    LEVELS = read_trace(sys.argv[2])
    TOP = {tuple(sorted(c)): c for c in nx.connected_components(PNET)}
    if "--clus" in sys.argv:
        assemble_clus(sys.argv[sys.argv.index("--clus") + 1], sys.argv[3])
        return
    out = open(sys.argv[3], "a")
    workers = int(sys.argv[sys.argv.index("--workers") + 1]) if "--workers" in sys.argv else 4
    import multiprocessing as mp
    if "--sequence" in sys.argv:
        global SEQ_LEVEL
        SEQ_LEVEL = int(sys.argv[sys.argv.index("--sequence") + 1])
        first_step = int(sys.argv[sys.argv.index("--from-step") + 1]) if "--from-step" in sys.argv else 0
        steps = list(range(first_step, len(LEVELS[SEQ_LEVEL][0]) + 1))   # the extra one checks the halves
        with mp.get_context("fork").Pool(workers) as pool:
            for r in pool.imap_unordered(check_step, steps):
                out.write(json.dumps(r) + "\n")
                out.flush()
        return
    max_nodes = int(sys.argv[sys.argv.index("--max-nodes") + 1]) if "--max-nodes" in sys.argv else 1 << 30
    todo = [i for i in range(len(LEVELS)) if len(LEVELS[i][1]) + len(LEVELS[i][2]) <= max_nodes]
    todo.sort(key=lambda i: -(len(LEVELS[i][1]) + len(LEVELS[i][2])) ** 2 * max(1, len(LEVELS[i][0])))   # longest first
    with mp.get_context("fork").Pool(workers) as pool:
        for r in pool.imap_unordered(check_level, todo):
            out.write(json.dumps(r) + "\n")
            out.flush()


if __name__ == "__main__":
    main()
