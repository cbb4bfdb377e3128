"""Minimal pytree utilities (≙ jax.tree_flatten/unflatten/map/leaves/structure used at
reference vkjax/function.py:27-28,40-42 and kompute_jaxpr_interpreter.py:69).

Containers: tuple, namedtuple, list, dict (sorted keys), None (an empty node, as in JAX).
If real JAX is importable its tree functions would work equally; these exist because it is not.
"""
import typing as tp


class PyTreeDef:
    """Immutable; the hash is computed once (a ResNet-50 state tree has ~340 nodes and is looked up on every call)."""
    __slots__ = ('kind', 'meta', 'children', 'num_leaves', '_hash')

    def __init__(self, kind, meta, children):
        self.kind = kind          # 'leaf' | 'none' | 'tuple' | 'list' | 'dict'
        self.meta = meta          # dict keys for 'dict'
        self.children = tuple(children)
        self.num_leaves = 1 if kind == 'leaf' else sum(c.num_leaves for c in self.children)
        self._hash = hash((kind, meta, tuple(c._hash for c in self.children)))

    def __eq__(self, other):
        if self is other:
            return True
        return (isinstance(other, PyTreeDef) and self._hash == other._hash and self.kind == other.kind
                and self.meta == other.meta and self.children == other.children)

    def __hash__(self):
        return self._hash

    def __repr__(self):
        if self.kind == 'leaf':
            return '*'
        if self.kind == 'none':
            return 'None'
        inner = ', '.join(map(repr, self.children))
        if self.kind == 'dict':
            inner = ', '.join(f'{k!r}: {c!r}' for k, c in zip(self.meta, self.children))
            return 'PyTreeDef({' + inner + '})'
        return f'PyTreeDef({self.kind}[{inner}])'


_LEAF = PyTreeDef('leaf', None, ())
_NONE = PyTreeDef('none', None, ())


def tree_flatten(tree) -> tp.Tuple[list, PyTreeDef]:
    leaves = []

    def rec(x):
        if x is None:
            return _NONE
        if isinstance(x, tuple) and hasattr(x, '_fields'):   # namedtuple: treat as tuple
            return PyTreeDef('tuple', type(x), [rec(c) for c in x])
        if isinstance(x, tuple):
            return PyTreeDef('tuple', None, [rec(c) for c in x])
        if isinstance(x, list):
            return PyTreeDef('list', None, [rec(c) for c in x])
        if isinstance(x, dict):
            keys = tuple(sorted(x.keys()))
            return PyTreeDef('dict', keys, [rec(x[k]) for k in keys])
        leaves.append(x)
        return _LEAF

    treedef = rec(tree)
    return leaves, treedef


def tree_unflatten(treedef: PyTreeDef, leaves):
    it = iter(leaves)

    def rec(td):
        if td.kind == 'leaf':
            return next(it)
        if td.kind == 'none':
            return None
        if td.kind == 'tuple':
            if td.meta is not None:                      # namedtuple: rebuilt as its own type
                return td.meta(*[rec(c) for c in td.children])
            return tuple(rec(c) for c in td.children)
        if td.kind == 'list':
            return [rec(c) for c in td.children]
        if td.kind == 'dict':
            return {k: rec(c) for k, c in zip(td.meta, td.children)}
        raise TypeError(td.kind)

    out = rec(treedef)
    return out


def tree_leaves(tree):
    return tree_flatten(tree)[0]


def tree_structure(tree):
    return tree_flatten(tree)[1]


def tree_map(f, tree, *rest):
    leaves, td = tree_flatten(tree)
    others = [td_flatten_up_to(td, r) for r in rest]
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


tree_multimap = tree_map


def td_flatten_up_to(td: PyTreeDef, tree):
    """Flatten `tree` only as deep as `td` goes (so leaves of td may be subtrees of tree)."""
    out = []

    def rec(t, x):
        if t.kind == 'leaf':
            out.append(x)
        elif t.kind == 'none':
            pass
        elif t.kind == 'dict':
            for k, c in zip(t.meta, t.children):
                rec(c, x[k])
        else:
            if len(x) != len(t.children):
                raise ValueError('pytree structure mismatch')
            for c, xi in zip(t.children, x):
                rec(c, xi)

    rec(td, tree)
    return out
