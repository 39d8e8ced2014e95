"""TEST INFRASTRUCTURE -- minimal `prettytable` stand-in (not installed in this image) so the
reference's LIRA_smallscale.py (:22) can be imported unmodified by oracle/make_golden.py."""


class PrettyTable:
    def __init__(self, field_names=None):
        self.field_names = list(field_names or [])
        self.float_format = ""
        self._rows = []

    def add_row(self, row):
        self._rows.append(list(row))

    def __str__(self):
        lines = [" | ".join(str(f) for f in self.field_names)]
        for r in self._rows:
            lines.append(" | ".join(f"{v:.4f}" if isinstance(v, float) else str(v) for v in r))
        return "\n".join(lines)
