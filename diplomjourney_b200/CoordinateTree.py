"""CoordinateTree with the reference's interface (CoordinateTree.py:4-36) and index rules,
without its storage cost.

The reference allocates ``np.empty(S + S**2 + S**3, tuple)`` (736 MB of None at S=451) and
fills 3*S slots.  Here the flat index space is kept -- layer offsets 0, S, S+S**2 and
``get_index_of_parent`` behave identically -- but slots live in a dict, so construction is
O(1).  On the GPU path the tree is never materialised at all: node k of layer d is the child
of node k of layer d-1 (math_model_tree.py:332,347-348), i.e. candidate k holds control k for
the whole horizon, and that is what the HELD kernel evaluates.
"""


class CoordinateTree:
    def __init__(self, size_max_1):
        self.size_1 = size_max_1
        self.size_2 = size_max_1 * size_max_1
        self.size_3 = size_max_1 * size_max_1 * size_max_1
        self.tree = {}

    def __str__(self):
        return str([self.tree.get(i) for i in sorted(self.tree)])

    def _check(self, index):
        if not -self.get_size() <= index < self.get_size():
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (index, self.get_size()))
        return index + self.get_size() if index < 0 else index

    def __getitem__(self, index):
        return self.tree.get(self._check(index))

    def __setitem__(self, index, coordinates):
        self.tree[self._check(index)] = coordinates

    def layer_of(self, index_of_element):
        if index_of_element < self.size_1:
            return 0
        return 1 if index_of_element < self.size_1 + self.size_2 else 2

    def get_index_of_parent(self, index_of_element):
        """Layer 0: the element itself; layer 1: its column in layer 0; layer 2:
        [parent in layer 1, grandparent in layer 0] (CoordinateTree.py:20-30)."""
        layer = self.layer_of(index_of_element)
        if layer == 0:
            return index_of_element
        if layer == 1:
            return (index_of_element - self.size_1) % self.size_1
        parent = self.size_1 + (index_of_element - self.size_1 - self.size_2) % self.size_1
        return [parent, self.get_index_of_parent(parent)]

    def get_size(self):
        return self.size_1 + self.size_2 + self.size_3

    def clear(self):
        self.tree = {}
