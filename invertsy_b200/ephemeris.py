"""
Sun position from place and time, vectorised: the host-side set-up that feeds
(theta_s, phi_s) into the sky of the rendering path.  It restates the NOAA solar-position spreadsheet
arithmetic that the reference's `Sun.compute` chains together scalar by scalar
(reference src/invertsy/env/ephemeris.py:79-124 and the helper formulas :231-753), in radians like the
reference, for whole arrays of dates / places at once -- a sweep over a day or a year is one call.

`Observer`, `get_seville_observer` keep the small surface `Sky.from_observer` needs
(reference src/invertsy/env/observer.py:22-173, :175-192, sky.py:580-611), including the reference's
habit of storing the Seville latitude / longitude *numbers* unconverted (observer.py:186-189).
"""
from datetime import datetime, timezone

import numpy as np


def _day_fraction_and_jd(dates):
    dates = np.atleast_1d(np.asarray(dates, dtype=object))
    frac = np.array([(d.hour + (d.minute + d.second / 60) / 60) / 24 for d in dates], dtype=np.float64)
    jd = np.array([d.toordinal() for d in dates], dtype=np.float64) + 1721424.5 + frac     # ephemeris.py:231-247
    return frac, jd


def sun_position(lon, lat, dates, tz=0.):
    """
    Altitude (refraction-corrected) and azimuth of the sun, radians, for longitude / latitude (radians),
    datetimes and time-zone offset(s) in hours.  Inputs broadcast; returns (alt, az) arrays.
    """
    frac, jd = _day_fraction_and_jd(dates)
    lon, lat, tz = (np.asarray(v, dtype=np.float64) for v in (lon, lat, tz))
    jc = (jd - 2451545) / 36525                                                              # :250-264
    # orbital elements
    mean_long = np.deg2rad((280.46646 + jc * (36000.76983 + jc * 0.0003032)) % 360)          # :267-281
    mean_anom = np.deg2rad(357.52911 + jc * (35999.05029 - 0.0001537 * jc))                  # :284-298
    ecc = 0.016708634 - jc * (0.000042037 + 0.0000001267 * jc)                               # :301-315
    centre = np.deg2rad(np.sin(mean_anom) * (1.914602 - jc * (0.004817 + 0.000014 * jc)) +
                        np.sin(2 * mean_anom) * (0.019993 - 0.000101 * jc) +
                        np.sin(3 * mean_anom) * 0.000289)                                    # :318-336
    true_long = mean_long + centre                                                           # :339
    omega = np.deg2rad(125.04 - 1934.136 * jc)
    app_long = true_long - np.deg2rad(0.00569 + 0.00478 * np.sin(omega))                     # :395-411
    mean_obliq = np.deg2rad(23 + (26 + (21.448 - jc * (46.815 + jc * (0.00059 - jc * 0.001813))) / 60) / 60)
    obliq = mean_obliq + np.deg2rad(0.00256) * np.cos(omega)                                 # :414-447
    decl = np.arcsin(np.sin(obliq) * np.sin(app_long))                                       # :469-485
    # equation of time (minutes) and true solar time
    y = np.square(np.tan(obliq / 2))                                                         # :488-502
    eot = 4 * np.rad2deg(y * np.sin(2 * mean_long) - 2 * ecc * np.sin(mean_anom) +
                         4 * ecc * y * np.sin(mean_anom) * np.cos(2 * mean_long) -
                         0.5 * np.square(y) * np.sin(4 * mean_long) -
                         1.25 * np.square(ecc) * np.sin(2 * mean_anom))                      # :505-531
    tst = (frac * 1440 + eot + 4 * np.rad2deg(lon) - 60 * tz) % 1440                         # :632-653
    ha = np.deg2rad(np.where(tst < 0, tst / 4 + 180, tst / 4 - 180))                         # :656-671
    # horizontal coordinates
    zen = np.arccos(np.sin(lat) * np.sin(decl) + np.cos(lat) * np.cos(decl) * np.cos(ha))    # :674-692
    elev = np.pi / 2 - zen                                                                   # :695
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t = np.tan(elev)
        deg = np.rad2deg(elev)
        refr = np.where(deg > 85, 0.,
               np.where(deg > 5, np.deg2rad((1 / t - 0.07 / np.power(t, 3) + 0.000086 / np.power(t, 5)) / 3600),
               np.where(deg > -0.575,
                        np.deg2rad((1735 + elev * (-518.2 - elev * (-518.2 + elev * (103.4 + elev * (-12.79 + elev * 0.711))))) / 3600),
                        np.deg2rad((-20.772 / t) / 3600))))                                  # :712-733
        alt = elev + refr                                                                    # :736
        core = np.arccos(((np.sin(lat) * np.cos(zen)) - np.sin(decl)) / (np.cos(lat) * np.sin(zen)))
    az = np.where(ha > 0, (core + np.pi) % (2 * np.pi), (np.deg2rad(540) - core) % (2 * np.pi))   # :739-753
    return alt, az


class Observer(object):
    """Place and time on Earth (reference observer.py:22-173): lon / lat in radians unless degrees=True."""

    def __init__(self, lon=None, lat=None, date=None, city=None, degrees=False):
        date = datetime.now() if date is None else date
        conv = np.deg2rad if degrees else (lambda v: v)
        self._lon = None if lon is None else conv(float(lon))
        self._lat = None if lat is None else conv(float(lat))
        self._city = city
        self._inrad = not degrees
        self.on_change = None
        self.date = date

    @property
    def date(self):
        return self._date

    @date.setter
    def date(self, value):
        self._date = value
        # hours ahead of GMT, as the reference derives it from the local clock (observer.py:48-52)
        self._tz = value.hour - value.astimezone(timezone.utc).hour
        if self.on_change is not None:
            self.on_change()

    def _get(self, v):
        return v if self._inrad else np.rad2deg(v)

    def _set(self, value):
        return float(value) if self._inrad else np.deg2rad(float(value))

    lon = property(lambda self: self._get(self._lon), lambda self, v: setattr(self, "_lon", self._set(v)))
    lat = property(lambda self: self._get(self._lat), lambda self, v: setattr(self, "_lat", self._set(v)))

    @property
    def tzgmt(self):
        return self._tz

    @property
    def city(self):
        return self._city


def get_seville_observer():
    """Seville, Spain (reference observer.py:175-192; the numbers are stored as given there)."""
    sev = Observer()
    sev.lat = '37.392509'
    sev.lon = '-5.983877'
    sev._city = "Seville"
    return sev


def sun_position_for_observer(obs=None, date=None):
    """(alt, az) for `Sky.from_observer` (reference sky.py:597-609): Seville, 21/06/2021 10:00 by default."""
    if obs is None:
        obs = get_seville_observer()
        obs.date = datetime(2021, 6, 21, 10, 0, 0) if date is None else date
    alt, az = sun_position(obs._lon, obs._lat, [obs.date], obs.tzgmt)
    return float(alt[0]), float(az[0])
