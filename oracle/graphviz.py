"""TEST INFRASTRUCTURE (oracle) — restatement of the reference's Graphviz dump, src/graphviz.rs:5-163.

Only tests/ may import this module; the product (kanpyo_b200/graphviz.py) never does.  It works on the
ORACLE's lattice (`OracleTokenizer.lattice`, the restated `Lattice::build` + `viterbi`), so the dot text
it produces is independent of the CUDA path and of the product's host code.  The reference holds no
golden dot output (nothing in `src/tests.rs` calls `graphviz`), so this side is pinned only by
following the Rust statement by statement; each block cites the lines it restates.

Rust's `BTreeSet<Node>` / `BTreeMap<&Node, _>` order and merge nodes by the derived `Ord` of
src/lattice/node.rs:6-24: enum variant first (Dummy < Known < Unknown), then the fields in
declaration order — Dummy {byte_pos, char_pos, morph}; Word {id, byte_pos, char_pos, morph, surface};
Morph {left_id, right_id, cost} (kanpyo-dict/src/morph.rs:7-11); `String` compares bytewise.  Python
tuples of ints / bytes compare the same way.
"""
from __future__ import annotations

from collections import deque

DUMMY, KNOWN, UNKNOWN = 0, 1, 2


def _rust_node(row, text: bytes):
    """One row of the oracle's node table (kind, id, byte_pos, char_pos, end_char, left, right, cost,
    byte_len) as a value ordered like the Rust `Node`."""
    kind, nid, byte_pos, char_pos, _end, left, right, cost, byte_len = (int(x) for x in row)
    morph = (left, right, cost)
    if kind == DUMMY:
        return (DUMMY, byte_pos, char_pos, morph)
    return (kind, nid, byte_pos, char_pos, morph, bytes(text[byte_pos:byte_pos + byte_len]))


class Graphviz:
    """`pub struct Graphviz<'a> { pub lattice: Lattice<'a> }` (src/graphviz.rs:5-7)."""

    def __init__(self, oracle_dict, lattice: dict, text: bytes):
        self.dict = oracle_dict
        self.nodes = [_rust_node(r, text) for r in lattice["nodes"]]
        off, idx = lattice["edge_off"], lattice["edge_idx"]
        self.edges = [[int(i) for i in idx[int(off[k]):int(off[k + 1])]] for k in range(len(off) - 1)]
        self.path = [int(i) for i in lattice["path"]]            # Lattice::viterbi()'s node list (indices)

    @staticmethod
    def _char_pos(node):                                          # Node::char_pos (node.rs:41-46)
        return node[2] if node[0] == DUMMY else node[3]

    @staticmethod
    def _morph(node):                                             # Node::morph (node.rs:48-53)
        return node[3] if node[0] == DUMMY else node[4]

    def bfs(self, start, bests):                                  # src/graphviz.rs:10-28
        visited = set()
        queue = deque([start])
        while queue:
            node = queue.popleft()
            if node in visited:                                   # `!visited.insert(node.clone())`
                continue
            visited.add(node)
            for i in self.edges[self._char_pos(node)]:
                cand = self.nodes[i]
                if cand in visited:
                    continue
                if cand[0] == UNKNOWN and cand not in bests:
                    continue
                queue.append(cand)
        return sorted(visited)                                    # `visited.into_iter().collect()`: BTreeSet order

    def graphviz(self, dpi: int, full_state: bool) -> str:        # src/graphviz.rs:30-163
        out = []
        bests = set(self.nodes[i] for i in self.path)             # :31-35
        out.append("graph lattice {")                            # :36-40
        out.append("dpi=%d;" % dpi)
        out.append("graph [style=filled, splines=true, overlap=false, fontsize=30, rankdir=LR]")
        out.append('edge [fontname=Helvetica, fontcolor=red, color="#606060"]')
        out.append('node [shape=box, style=filled, fillcolor="#e8e8f0", fontname=Helvetica]')
        if not full_state:                                        # :42-54
            visible_nodes = self.bfs(self.nodes[-1], bests)
        else:
            visible_nodes = list(self.nodes)
        for visible_id, node in enumerate(visible_nodes):         # :56-121
            if node[0] == KNOWN:
                rows, names = self.dict.features
            elif node[0] == UNKNOWN:
                rows, names = self.dict.unk_features
            if node[0] == DUMMY:
                label = "BOS" if visible_id == 0 else "EOS"
            else:
                feats = [names[k] for k in rows[node[1] - 1]]
                label = "%s\n%s\n%d" % (node[5].decode("utf-8"), "/".join(f for f in feats if f != "*"), node[4][2])
            color = {KNOWN: "black", UNKNOWN: "red", DUMMY: "blue"}[node[0]]
            if node in bests or node[0] == DUMMY:
                out.append('%d [label="%s", shape=ellipse, color=%s, peripheries=2]' % (visible_id, label, color))
            else:
                shape = {KNOWN: "box", UNKNOWN: "diamond", DUMMY: "ellipse"}[node[0]]
                out.append('%d [label="%s", shape=%s, color=%s]' % (visible_id, label, shape, color))
        node_to_visible_id = {}                                   # :122-126 (collect into a map: later entries replace)
        for vid, node in enumerate(visible_nodes):
            node_to_visible_id[node] = vid
        conn, row = self.dict.conn, int(self.dict.conn_row)
        for edge in self.edges:                                   # :127-161
            for i in edge:
                node = self.nodes[i]
                if node not in node_to_visible_id:
                    continue
                vid = node_to_visible_id[node]
                for j in self.edges[self._char_pos(node)]:
                    from_node = self.nodes[j]
                    if from_node not in node_to_visible_id:
                        continue
                    from_id = node_to_visible_id[from_node]
                    if from_id == vid:
                        continue
                    # ConnectionTable::get(row = from.right_id, col = node.left_id) = data[self.row*col + row]
                    label = int(conn[row * self._morph(node)[0] + self._morph(from_node)[1]])
                    ok1 = from_node in bests or from_node[0] == DUMMY
                    ok2 = node in bests or node[0] == DUMMY
                    if ok1 and ok2:
                        out.append('%d -- %d [label="%d", style=bold, color=blue, fontcolor=blue]' % (from_id, vid, label))
                    else:
                        out.append('%d -- %d [label="%d"]' % (from_id, vid, label))
        out.append("}")                                           # :162
        return "\n".join(out) + "\n"


def graphviz(oracle_tokenizer, text: str, dpi: int = 48, full_state: bool = False) -> str:
    """What `kanpyo graphviz` prints for `text` (src/bin/kanpyo.rs:127-148 -> Graphviz::graphviz)."""
    raw = text.encode("utf-8")
    return Graphviz(oracle_tokenizer.d, oracle_tokenizer.lattice(text), raw).graphviz(dpi, full_state)
