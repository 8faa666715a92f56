"""Small helpers to edit <flame> attributes in genome XML text (tests, bench, tools).

The parser that gives these attributes meaning is the reference's own (parser.c:333-493);
here they are only rewritten textually, e.g. size / scale / quality / supersample /
temporal_samples / estimator_radius, the way an artist edits a .flam3 file.
"""
import re


def set_flame_attrs(xml, **attrs):
    """Set/replace attributes on every <flame ...> opening tag."""

    def fix(m):
        tag = m.group(0)
        for k, v in attrs.items():
            pat = re.compile(r'\s%s="[^"]*"' % re.escape(k))
            tag = pat.sub("", tag)
            tag = tag[:-1] + ' %s="%s">' % (k, v)
        return tag

    return re.sub(r"<flame\b[^>]*>", fix, xml)


def resize(xml, width, height, scale=None, quality=None, oversample=None, **extra):
    attrs = {"size": "%d %d" % (width, height)}
    if scale is not None:
        attrs["scale"] = repr(float(scale))
    if quality is not None:
        attrs["quality"] = repr(float(quality))
    if oversample is not None:
        attrs["supersample"] = str(int(oversample))
    attrs.update(extra)
    return set_flame_attrs(xml, **attrs)


def flame_attr(xml, name, default=None):
    """Value of attribute `name` on the first <flame ...> tag."""
    m = re.search(r"<flame\b[^>]*?\s%s=\"([^\"]*)\"" % re.escape(name), xml)
    return m.group(1) if m else default


def rescale(xml, width, height, quality=None, **extra):
    """Resize keeping the field of view: `scale` (pixels per unit) follows the width."""
    w0 = float(flame_attr(xml, "size", "%d %d" % (width, height)).split()[0])
    s0 = float(flame_attr(xml, "scale", "1"))
    return resize(xml, width, height, scale=s0 * width / w0, quality=quality, **extra)
