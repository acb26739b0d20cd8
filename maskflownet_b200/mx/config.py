"""Null-safe configuration reader with the interface the reference's models expect
(`config.network.flow_multiplier.get(1.)`, `.value`) -- mirrors network/config/__init__.py:1-22 of the reference."""
from __future__ import annotations

import logging

_log = logging.getLogger("maskflownet_b200.config")


class Reader:
    __slots__ = ("_node", "_path")

    def __init__(self, node=None, path=""):
        object.__setattr__(self, "_node", node)
        object.__setattr__(self, "_path", path)

    def __getattr__(self, key):
        node = object.__getattribute__(self, "_node")
        child = node.get(key) if isinstance(node, dict) else None
        return Reader(child, f"{object.__getattribute__(self, '_path')}.{key}")

    def get(self, default=None):
        node = object.__getattribute__(self, "_node")
        if node is None:
            _log.debug("config%s not set, using %r", object.__getattribute__(self, "_path"), default)
            return default
        return node

    @property
    def value(self):
        return object.__getattribute__(self, "_node")
