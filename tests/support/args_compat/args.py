"""Stand-in for the reference's flag module (args.py, 38 lines: argparse + defaults + derived
flags) used ONLY when the reference tree is not on sys.path (e.g. on the GPU box).  Same public
names: FLAGS, PARSER, add_argument, add_derivation, parse_flags, update_flags, apply_derivations."""
import argparse


class _Flags(argparse.Namespace):
    def __getattr__(self, name):  # unset flags fall back to their registered default (args.py:8-12)
        defaults = PARSER.defaults
        if name in defaults:
            return defaults[name]
        raise AttributeError(name)


PARSER = argparse.ArgumentParser()
PARSER.defaults = {}
PARSER.derivations = []
FLAGS = _Flags()


def add_argument(name, default, **kwargs):
    if kwargs.get("type") is bool:
        kwargs.update(nargs="?", const=True)
    try:
        PARSER.add_argument(name, **kwargs)
    except argparse.ArgumentError:
        pass  # registered twice (drop-in and launcher both declare it)
    PARSER.defaults[name.replace("-", "")] = default


def add_derivation(fn):
    PARSER.derivations.append(fn)


def apply_derivations(parser=PARSER):
    for _ in range(10):
        before = dict(FLAGS.__dict__)
        for fn in parser.derivations:
            fn()
        if FLAGS.__dict__ == before:
            return
    raise Exception("Could not find settings fixed point")


def parse_flags(argv=None):
    PARSER.parse_args(argv, namespace=FLAGS)
    apply_derivations(PARSER)


def update_flags(**kwargs):
    FLAGS.__dict__.update(**kwargs)
    apply_derivations(PARSER)
