"""Import the LIVE reference classes from /root/reference/src (build container
only -- the GPU box has no /root/reference).  TEST INFRASTRUCTURE ONLY.

Two in-memory shims make the reference importable without writing to its tree
(SURVEY.md §8c / §9.1): ``colorlog`` (not installed) and ``config.mypath``
(git-ignored in the reference, ``config/mypath.py.example:1``)."""
from __future__ import annotations

import logging
import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(REF_SRC)


def load():
    """Returns a namespace with OSVOS_VGG, layers module, providers, settings."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_SRC)
    if "colorlog" not in sys.modules:
        cl = types.ModuleType("colorlog")

        class ColoredFormatter(logging.Formatter):
            def __init__(self, fmt=None, *a, **k):
                k.pop("log_colors", None)
                k.pop("reset", None)
                k.pop("secondary_log_colors", None)
                k.pop("style", None)
                super().__init__("%(levelname)s %(message)s")

        cl.ColoredFormatter = ColoredFormatter
        cl.getLogger = logging.getLogger
        cl.StreamHandler = logging.StreamHandler
        sys.modules["colorlog"] = cl
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import config  # noqa: F401  (reference package)
    if "config.mypath" not in sys.modules:
        from config.path_abstract import PathAbstract
        mp = types.ModuleType("config.mypath")
        names = [n for n in dir(PathAbstract) if not n.startswith("_")]
        body = {}
        for n in names:
            if n.startswith("is_"):
                body[n] = staticmethod(lambda: False)
            else:
                body[n] = staticmethod(lambda: "/nonexistent")
        mp.Path = type("Path", (PathAbstract,), body)
        sys.modules["config.mypath"] = mp
        config.mypath = mp
    logging.disable(logging.INFO)
    from networks.osvos_vgg import OSVOS_VGG
    import layers.osvos_layers as L
    from util.network_provider import VGGOnlineProvider, VGGOfflineProvider
    from util.settings import OnlineSettings, OfflineSettings
    ns = types.SimpleNamespace(OSVOS_VGG=OSVOS_VGG, layers=L, VGGOnlineProvider=VGGOnlineProvider,
                               VGGOfflineProvider=VGGOfflineProvider, OnlineSettings=OnlineSettings,
                               OfflineSettings=OfflineSettings)
    return ns
