"""Tolerant reader for the reference's .ini files (IniFiles/*.ini).

The reference compares raw strings including their quotes (`== "'yes'"`, RKD2Q9.py:36) and exits on any
missing key; several keys it asks for do not exist in its own shipped files (`SurfaceTensionValue` vs
`SurfaceTension`, `[BodyForce] isBodyForce`, RKD2Q9.py:72,197 vs RKtwophasesetup2D.ini:14,37).  This reader
accepts both spellings, quoted or unquoted values, is case-insensitive and takes defaults."""
import configparser
import os


class IniError(SystemExit):
    """the reference's error behaviour is print + sys.exit(); IniError is a SystemExit with the message"""


class Ini:
    def __init__(self, directory, filename):
        self.path = os.path.join(directory, filename)
        if not os.path.exists(self.path):
            raise IniError("Could not find %s" % self.path)
        self.cfg = configparser.ConfigParser(inline_comment_prefixes=(";;",))
        self.cfg.read(self.path)
        self._sections = {s.lower(): s for s in self.cfg.sections()}

    def raw(self, section, *keys, default=None):
        sec = self._sections.get(section.lower())
        if sec is not None:
            for k in keys:
                if self.cfg.has_option(sec, k):
                    return self.cfg.get(sec, k).strip()
        if default is None:
            raise IniError("Could not find [%s] %s in %s, please check .ini file" % (section, "/".join(keys), self.path))
        return default

    def text(self, section, *keys, default=None):
        """value without its quotes: 'MRT' -> MRT"""
        return str(self.raw(section, *keys, default=default)).strip().strip("'\"")

    def quoted(self, section, *keys, default=None):
        """value in the reference's own convention, e.g. "'MRT'" (for attributes other code compares)"""
        return "'%s'" % self.text(section, *keys, default=default)

    def number(self, section, *keys, default=None, cast=float):
        return cast(float(self.text(section, *keys, default=None if default is None else str(default))))

    def integer(self, section, *keys, default=None):
        return self.number(section, *keys, default=default, cast=int)

    def numbers(self, section, *keys, default=None):
        raw = self.text(section, *keys, default=default)
        return [float(x) for x in raw.split(",") if x.strip()]

    def has_section(self, section):
        return section.lower() in self._sections
