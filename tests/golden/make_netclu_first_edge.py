"""For a network whose Girvan-Newman split the reference's netclu_ng.py cannot finish (mycoplasma64: one component of
11,325 genes; one networkx betweenness pass takes tens of minutes, the split needs 1,498), pins the FIRST step of the
script on its largest component: the graph is built as netclu_ng.py:43-56 builds it, the component taken as :149 takes
it, the working copy made as networkx' girvan_newman makes it, and the edge that most_valuable_edge returns is printed —
netclu_cc -g with PD_NETCLU_TRACE=2 prints the edges it removes (`rm a b`), the first of that component must be this one.

    python tests/golden/make_netclu_first_edge.py in.net      (needs networkx; prints "first a b betweenness")
"""
import sys

import networkx as nx


def main(path):
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
    coco = max(nx.connected_components(pnet), key=len)
    snet = pnet.subgraph(coco)
    g = snet.copy().to_undirected()
    bet = nx.edge_betweenness_centrality(g)
    e = max(bet, key=bet.get)
    print("first", e[0], e[1], repr(bet[e]), "nodes", g.number_of_nodes(), "edges", g.number_of_edges())


if __name__ == "__main__":
    main(sys.argv[1])
