"""The define-by-run graph: which Node stands for which Tensor (public surface of neograd/autograd/graph.py: `track`,
`nodes_dict`, add_edge / add_node / add_tensor / get_node / remove_tensor, reset_visited / reset_graph / zero_grad, and
the class attribute `graph` that `new_graph()` swaps).  Host bookkeeping only -- O(#ops) dictionary work per step."""
from .node import Node


class Graph:
  graph = None   # the Graph a `with new_graph():` block has put in force, else None (the global one is used)

  def __init__(self):
    self.track = True        # False inside no_track() / model.eval(): ops then create no nodes
    self.nodes_dict = {}     # Tensor -> Node

  # ---------------------------------------------------------------- construction
  def _node_of(self, tens):
    """The tensor's node, created on first sight (operands that are leaves enter the graph this way)."""
    node = self.nodes_dict.get(tens)
    if node is None:
      node = Node(tens)
      self.nodes_dict[tens] = node
    return node

  def add_node(self, node):
    self.nodes_dict[node.tens] = node

  def add_tensor(self, tens):
    self.nodes_dict[tens] = Node(tens)

  def add_edge(self, result_node, operands):
    """Registers an op's result and links it to the nodes of its operands (in operand order)."""
    self.add_node(result_node)
    for tens in operands:
      operand_node = self._node_of(tens)
      operand_node.add_child(result_node)
      result_node.add_parent(operand_node)

  def get_node(self, tens):
    return self.nodes_dict.get(tens)

  # ---------------------------------------------------------------- teardown
  @staticmethod
  def _unlink(node):
    # parent <-> child lists are reference cycles; cutting them lets a pass's intermediates (and any deferred arrays they
    # hold) die by reference count at once instead of waiting for the cyclic collector
    node.parents, node.children = [], []

  def remove_tensor(self, tens):
    """Drops a tensor whose parents have all been visited (KeyError if it is not in the graph, as in the reference)."""
    self._unlink(self.nodes_dict.pop(tens))

  def reset_graph(self):
    old, self.nodes_dict = self.nodes_dict, {}
    for node in old.values():
      self._unlink(node)

  def reset_visited(self):
    for node in self.nodes_dict.values():
      node.visited = False

  def zero_grad(self):
    """Optimizer.zero_grad(all_members=True): every tensor currently in the graph."""
    for tens in self.nodes_dict:
      tens.zero_grad()

  def __repr__(self):
    return 'Graph()'

  def __str__(self):
    return f'Graph( {self.nodes_dict} )'
