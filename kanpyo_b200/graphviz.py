"""Graphviz dump of a lattice built on the device — host-side mirror of src/graphviz.rs:10-163
(`kanpyo graphviz`).  Debug visualisation over `Tokenizer.lattice()` (kp_lattice_dump); nothing here
is on the hot path.

Node identity and order follow the reference's derived `Ord` on `Node` (src/lattice/node.rs:7-24):
variant (Dummy < Known < Unknown), then for Dummy (byte_pos, char_pos, morph), for a word
(id, byte_pos, char_pos, morph, surface); `BTreeSet<Node>` therefore merges equal nodes and numbers
the visible ones in that order.
"""
from __future__ import annotations

from collections import deque

import numpy as np


def _key(nd, surface: bytes):
    morph = (int(nd["left_id"]), int(nd["right_id"]), int(nd["cost"]))
    if nd["cls"] == 0:
        return (0, int(nd["byte_pos"]), int(nd["char_pos"]), morph)
    return (int(nd["cls"]), int(nd["id"]), int(nd["byte_pos"]), int(nd["char_pos"]), morph, surface)


def graphviz(tokenizer, text: str, dpi: int = 48, full_state: bool = False) -> str:
    """The text `Graphviz { lattice }.graphviz(dpi, full_state)` prints (src/graphviz.rs:30-163)."""
    nodes = tokenizer.lattice(text)
    raw = text.encode("utf-8")
    char_byte = [i for i, b in enumerate(raw) if (b & 0xC0) != 0x80] + [len(raw)]
    n_chars = len(char_byte) - 1
    surf = [b"" if nd["cls"] == 0 else raw[int(nd["byte_pos"]):char_byte[int(nd["end_char"])]] for nd in nodes]
    keys = [_key(nd, s) for nd, s in zip(nodes, surf)]
    edges = [[] for _ in range(n_chars + 2)]                       # Lattice.edges: nodes by END char position
    for i, nd in enumerate(nodes):
        edges[int(nd["end_char"])].append(i)
    # best path = viterbi() (src/lattice.rs:144-153): follow pre from the last node; BOS is not on it
    bests, pos = set(), len(nodes) - 1
    while nodes[pos]["pre"] >= 0:
        bests.add(keys[pos])
        pos = int(nodes[pos]["pre"])
    if full_state:
        visible = list(range(len(nodes)))                          # `self.lattice.nodes.clone()`: insertion order
        vis_keys = [keys[i] for i in visible]
    else:                                                          # bfs from the last node (src/graphviz.rs:10-28)
        rep, queue = {}, deque([len(nodes) - 1])
        while queue:
            i = queue.popleft()
            if keys[i] in rep:
                continue
            rep[keys[i]] = i
            for j in edges[int(nodes[i]["char_pos"])]:
                if keys[j] in rep:
                    continue
                if nodes[j]["cls"] == 2 and keys[j] not in bests:  # unknown nodes only when on the best path
                    continue
                queue.append(j)
        vis_keys = sorted(rep)                                     # BTreeSet iteration order
        visible = [rep[k] for k in vis_keys]
    out = ["graph lattice {", "dpi=%d;" % dpi,
           "graph [style=filled, splines=true, overlap=false, fontsize=30, rankdir=LR]",
           'edge [fontname=Helvetica, fontcolor=red, color="#606060"]',
           'node [shape=box, style=filled, fillcolor="#e8e8f0", fontname=Helvetica]']
    d = tokenizer.dict
    for vid, i in enumerate(visible):
        nd = nodes[i]
        cls = int(nd["cls"])
        if cls == 0:
            label = "BOS" if vid == 0 else "EOS"
        else:
            rows, names = d.features if cls == 1 else d.unk_features
            feats = "/".join(s for s in (names[k] for k in rows[int(nd["id"]) - 1]) if s != "*")
            label = "%s\n%s\n%d" % (surf[i].decode("utf-8"), feats, int(nd["cost"]))
        color = ("blue", "black", "red")[cls]
        if keys[i] in bests or cls == 0:
            out.append('%d [label="%s", shape=ellipse, color=%s, peripheries=2]' % (vid, label, color))
        else:
            out.append('%d [label="%s", shape=%s, color=%s]' % (vid, label, ("ellipse", "box", "diamond")[cls], color))
    vid_of = {}
    for vid, k in enumerate(vis_keys):
        vid_of[k] = vid                                            # BTreeMap<&Node, id>: the last equal node wins
    conn, row = np.asarray(d.conn), int(d.conn_row)
    for bucket in edges:
        for i in bucket:
            if keys[i] not in vid_of:
                continue
            to_id, to = vid_of[keys[i]], nodes[i]
            for j in edges[int(to["char_pos"])]:
                if keys[j] not in vid_of:
                    continue
                from_id, frm = vid_of[keys[j]], nodes[j]
                if from_id == to_id:
                    continue
                label = int(conn[row * int(to["left_id"]) + int(frm["right_id"])])      # connection.rs:12-14
                ok1 = keys[j] in bests or frm["cls"] == 0
                ok2 = keys[i] in bests or to["cls"] == 0
                if ok1 and ok2:
                    out.append('%d -- %d [label="%d", style=bold, color=blue, fontcolor=blue]' % (from_id, to_id, label))
                else:
                    out.append('%d -- %d [label="%d"]' % (from_id, to_id, label))
    out.append("}")
    return "\n".join(out) + "\n"
