"""Stub of `sorcery` (imported by pyrl.utils.meta.magic_utils, never used on the hot path).
Test infrastructure only."""


def _unavailable(*args, **kwargs):
    raise RuntimeError("sorcery is stubbed out in the oracle shim")


assigned_names = unpack_keys = unpack_attrs = dict_of = print_args = _unavailable
call_with_name = delegate_to_attr = maybe = select_from = _unavailable
