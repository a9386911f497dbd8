"""The global define-by-run graph: Tensor -> Node map (reference: neograd/autograd/graph.py)."""
from .node import Node


class Graph:
  graph = None   # the graph currently in force when a `new_graph()` block is active (graph.py:21)

  def __init__(self):
    self.nodes_dict = {}
    self.track = True

  def add_edge(self, result_node, operands):
    self.add_node(result_node)
    for operand in operands:
      node = self.nodes_dict.get(operand)
      if node is None:
        node = self.nodes_dict[operand] = Node(operand)
      result_node.add_parent(node)
      node.add_child(result_node)

  def add_node(self, node):
    self.nodes_dict[node.tens] = node

  def get_node(self, tens):
    return self.nodes_dict.get(tens)

  def add_tensor(self, tens):
    self.nodes_dict[tens] = Node(tens)

  def remove_tensor(self, tens):
    node = self.nodes_dict.pop(tens)   # KeyError on double removal, as in the reference
    # a node is removed once all its parents were visited: nothing walks its edges again.  Dropping them breaks the
    # parent <-> child reference cycles, so the pass's intermediate tensors (and deferred arrays) die by reference count
    # right here instead of waiting for the garbage collector.
    node.parents = []
    node.children = []

  def reset_visited(self):
    for node in self.nodes_dict.values():
      node.visited = False

  def reset_graph(self):
    for node in self.nodes_dict.values():
      node.parents = []
      node.children = []
    self.nodes_dict = {}

  def zero_grad(self):
    for tens in self.nodes_dict:
      tens.zero_grad()

  def __repr__(self):
    return 'Graph()'

  def __str__(self):
    return f'Graph( {self.nodes_dict} )'
