"""Region / mutation value types with the reference's attribute names and text format
(poreseq/Util.py:1-111), so code written against `poreseq.Util` keeps working."""


class RegionInfo(object):
    """'name', 'start:end' or 'name:start:end' (poreseq/Util.py:1-30)."""

    def __init__(self, region=None):
        self.start = None
        self.end = None
        self.name = None
        if region is None:
            return
        parts = region.split(":")
        if len(parts) != 2:
            self.name = parts[0]
        if len(parts) > 1:
            self.start = int(parts[-2])
            self.end = int(parts[-1])


def _dot(s):
    return s if len(s) else "."


class MutationInfo(object):
    """start (0-based), orig, mut; '' (printed '.') for insertions/deletions (poreseq/Util.py:32-80)."""

    def __init__(self, info=None):
        self.start = 0
        self.orig = ""
        self.mut = ""
        if info is None:
            return
        if len(info) == 0 or info[0] == "#":
            self.start = -1
            return
        fields = info.split()
        if len(fields) != 3:
            self.start = -1
            return
        self.start = int(fields[0])
        self.orig = "" if fields[1] == "." else fields[1]
        self.mut = "" if fields[2] == "." else fields[2]

    def __str__(self):
        return "{}\t{}\t{}".format(self.start, _dot(self.orig), _dot(self.mut))


class MutationScore(object):
    """MutationInfo plus score (poreseq/Util.py:82-111)."""

    def __init__(self):
        self.start = 0
        self.orig = ""
        self.mut = ""
        self.score = 0

    def __str__(self):
        return "{}\t{}\t{}\t{}".format(self.start, _dot(self.orig), _dot(self.mut), self.score)
