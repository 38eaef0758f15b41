"""Plain data holders named after the reference's elemutils.py (which does not parse and is imported by
nothing, /root/reference/elemutils.py:1-11); kept for name parity only -- not on the hot path."""


class Node:
    def __init__(self, num, x, y):
        self.nodeNumber = num
        self.initialCoordinates = (x, y)


class Element:
    def __init__(self, num):
        self.elemNumber = num
        self.nodes = ()

    def setupConectivity(self, nodes):
        self.nodes = tuple(nodes)
