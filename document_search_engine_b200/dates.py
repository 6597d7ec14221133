"""Date ranges on the posting-list path (host side only).

The reference's search form documents ``date:[oct 1970 to dec 8 1970]`` and ``hospital date:"feb 1964"``
(``templates/search-form.html:26``, ``:39``); ``my_flask.py:189-193`` installs Whoosh's ``DateParserPlugin`` for them.
Whoosh indexes a DATETIME field as tiered numeric terms and compiles a ``DateRange`` (a ``NumericRange``) into the
``Or`` of the tier terms that cover the range, wrapped in a constant score: a matching document adds the query's
boost (1.0) to its score.  The flat index does the same with three readable tiers per document date -

    ``Y1970``   ``M1970-10``   ``D1970-10-08``

- in a field that is not scorable (W15: a posting scores its weight, 1, times the leaf boost), so a date range is an
OR-group of a few dozen ordinary posting lists and runs on the same kernels as any other query.

Resolution is one day: document dates must be dates (or datetimes at midnight), and a range
``[start, end]`` selects the days ``d`` with ``start <= d 00:00 <= end`` - exactly the documents Whoosh's numeric
range would select for such dates.
"""
import calendar
import re
from datetime import date, datetime, timedelta
from typing import List, Optional, Tuple

MONTHS = {m.lower(): i for i, m in enumerate(calendar.month_name) if m}
MONTHS.update({m.lower(): i for i, m in enumerate(calendar.month_abbr) if m})
MONTHS["sept"] = 9


def as_day(value) -> date:
    """The day a document is filed under; refuses times of day (the tiers stop at days)."""
    if isinstance(value, datetime):
        if (value.hour, value.minute, value.second, value.microsecond) != (0, 0, 0, 0):
            raise ValueError("dates are indexed at day resolution: %r has a time of day" % (value,))
        return value.date()
    if isinstance(value, date):
        return value
    raise TypeError("%r is not a date" % (value,))


def tier_tokens(value) -> Tuple[str, str, str]:
    d = as_day(value)
    return "Y%04d" % d.year, "M%04d-%02d" % (d.year, d.month), "D%04d-%02d-%02d" % (d.year, d.month, d.day)


def first_day(start) -> date:
    """First day ``d`` with ``start <= d 00:00``."""
    if isinstance(start, datetime):
        d = start.date()
        return d if start == datetime(d.year, d.month, d.day) else d + timedelta(days=1)
    return start


def last_day(end) -> date:
    """Last day ``d`` with ``d 00:00 <= end``."""
    return end.date() if isinstance(end, datetime) else end


def range_cover(lo: date, hi: date) -> List[str]:
    """Tier tokens whose documents are exactly those dated ``lo .. hi`` (inclusive): whole years where a year fits,
    whole months where a month fits, single days at the ragged ends."""
    out: List[str] = []
    d = lo
    while d <= hi:
        if d.month == 1 and d.day == 1 and date(d.year, 12, 31) <= hi:
            out.append("Y%04d" % d.year)
            d = date(d.year + 1, 1, 1) if d.year < 9999 else hi + timedelta(days=1)
            continue
        mlast = date(d.year, d.month, calendar.monthrange(d.year, d.month)[1])
        if d.day == 1 and mlast <= hi:
            out.append("M%04d-%02d" % (d.year, d.month))
            d = mlast + timedelta(days=1)
            continue
        out.append("D%04d-%02d-%02d" % (d.year, d.month, d.day))
        d += timedelta(days=1)
    return out


class DateParseError(ValueError):
    """An unreadable date expression (the reference catches the parser's error and redirects, ``my_flask.py:193-196``)."""


_NUM = re.compile(r"^\d+$")


def parse_span(text: str) -> Tuple[datetime, datetime]:
    """``(first instant, last instant)`` of what ``text`` names: ``1964``, ``feb 1964``, ``dec 8 1970``, ``8 dec 1970``,
    ``1970-12``, ``1970-12-08``, ``19701208`` (the forms the reference's help shows, and ISO)."""
    t = text.strip().lower().replace(",", " ")
    m = re.match(r"^(\d{4})-(\d{1,2})(?:-(\d{1,2}))?$", t)
    if m:
        y, mo, dd = int(m.group(1)), int(m.group(2)), m.group(3)
        return _span(y, mo, int(dd) if dd else None)
    m = re.match(r"^(\d{4})(\d{2})(\d{2})$", t)
    if m:
        return _span(int(m.group(1)), int(m.group(2)), int(m.group(3)))
    words = t.split()
    year = month = day = None
    for w in words:
        if w in MONTHS and month is None:
            month = MONTHS[w]
        elif _NUM.match(w) and len(w) == 4 and year is None:
            year = int(w)
        elif _NUM.match(w) and len(w) <= 2 and day is None:
            day = int(w)
        else:
            raise DateParseError("cannot read the date %r" % text)
    if year is None or (day is not None and month is None):
        raise DateParseError("cannot read the date %r" % text)
    return _span(year, month, day)


def _span(year: int, month: Optional[int], day: Optional[int]) -> Tuple[datetime, datetime]:
    try:
        if month is None:
            a, b = date(year, 1, 1), date(year, 12, 31)
        elif day is None:
            a, b = date(year, month, 1), date(year, month, calendar.monthrange(year, month)[1])
        else:
            a = b = date(year, month, day)
    except ValueError as e:
        raise DateParseError(str(e))
    return datetime(a.year, a.month, a.day), datetime(b.year, b.month, b.day, 23, 59, 59, 999999)


def parse_range(text: str) -> Tuple[Optional[datetime], Optional[datetime]]:
    """``a to b`` (either side may be empty: open) or a single date expression."""
    m = re.match(r"^(.*?)\bto\b(.*)$", text.strip(), re.IGNORECASE)
    if not m:
        return parse_span(text)
    a, b = m.group(1).strip(), m.group(2).strip()
    return (parse_span(a)[0] if a else None), (parse_span(b)[1] if b else None)
