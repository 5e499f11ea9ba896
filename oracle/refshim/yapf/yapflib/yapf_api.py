"""Stub: the reference only calls FormatCode for Config.pretty_text. Test infrastructure only."""


def FormatCode(source, **kwargs):
    return source, False
