"""Occupancy-grid mapping on the GPU behind the reference's class API.

Drop-in for `from mapping import Mapping` (W12 w12-mapping / w12-mapping-online
course_agv_slam/scripts/mapping.py).  The device owns two int32 planes (endpoint hits,
traversals) for the life of the object; `pmap` / `datamap` are materialised from them.
"""
import ctypes

import numpy as np

from b2slam import _lib


class MapTicket(object):
    """Handle of a step submitted with Mapping.submit_scans / submit_batch."""

    def __init__(self, owner, ticket, out):
        self._owner, self._ticket, self._out = owner, ticket, out

    def wait(self):
        """Block until the step's occupancy is in host memory and return it (int8 (xw, yw), page-locked, overwritten
        by the second-next submit).  A step holding a coordinate the reference's int() raises on raises here
        (ValueError / OverflowError) and has been taken back out of the counts."""
        m = self._owner
        rc = m._L.b2s_mapping_wait(m._h, self._ticket)
        m._raise_nonfinite(rc)
        _lib.check(rc)
        m._host_map_blank = False
        return self._out


class Mapping(object):
    """[MAP]:7-51.  Mapping(xw, yw, xyreso) as in the reference, plus keyword weights:

    hit_weight 20.0 ([MAP]:45; use 4.0 for w12-mapping-online, [MAPO]:46), miss_weight 0.01
    ([MAP]:43), occ_threshold 10.0 ([MAP]:47).  The world->cell transform generalises the
    reference's literals: cells_per_m = 1/xyreso, offset = extent/2 (exactly 10, 10 for the
    reference's 200 x 200 x 0.1 m map).
    """

    def __init__(self, xw, yw, xyreso, hit_weight=20.0, miss_weight=0.01, occ_threshold=10.0,
                 device=-1):
        self.xw = int(xw)
        self.yw = int(yw)
        self.xyreso = float(xyreso)
        self.width_x = self.xw * self.xyreso
        self.width_y = self.yw * self.xyreso
        self.minx = -self.width_x / 2.0
        self.maxx = self.width_x / 2.0
        self.miny = -self.width_y / 2.0
        self.maxy = self.width_y / 2.0
        self.hit_weight = float(hit_weight)
        self.miss_weight = float(miss_weight)
        self.occ_threshold = float(occ_threshold)
        self._L = _lib.lib()
        _lib.require_device()
        h = ctypes.c_void_p()
        _lib.check(self._L.b2s_mapping_create(ctypes.byref(h), self.xw, self.yw, self.xyreso,
                                              self.hit_weight, self.miss_weight,
                                              self.occ_threshold, int(device)))
        self._h = h
        self._pmap8 = _lib.pinned_empty((self.xw, self.yw), np.int8)  # page-locked result buffer
        self._pmap8.fill(50)  # [MAP]:14 unknown = 50
        self._pmap64 = None            # float64 mirror handed out by update(), patched tile by tile
        self._tiles = np.empty(1024, dtype=np.int32)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._L.b2s_mapping_destroy(h)
            self._h = None

    @property
    def pmap(self):
        """[MAP]:14 occupancy as the reference's float64 (xw, yw) array in {0, 50, 100}: the live array
        update() maintains (like the reference's), rebuilt here if a batched call moved past it."""
        if self._pmap64 is None:
            self._pmap64 = self._host_map().astype(np.float64)
        return self._pmap64

    def _host_map(self):
        """The int8 host copy of the occupancy, made consistent with a reset that nobody read back."""
        if getattr(self, "_host_map_blank", False):
            self._pmap8.fill(50)
            self._host_map_blank = False
        return self._pmap8

    # ------------------------------------------------------------------ reference methods

    def update(self, ox, oy, center_x, center_y):
        """[MAP]:22-51.  ox, oy (N,) world-frame endpoints, center_* scalar (or 1-element array,
        as slam_ekf.py:90 passes).  Returns the (xw, yw) float64 occupancy in {0, 50, 100}.

        Coordinates cross the boundary as float64, the dtype the reference evaluates int(10*(v+10)) on
        ([MAP]:33-36; slam_ekf.py:89-90 passes float64 obs and xEst), so a coordinate on or next to a cell
        boundary lands in the reference's cell.  Raises ValueError / OverflowError on NaN / inf where the
        reference's int() does (infinite ox alone is skipped, [MAP]:30) -- but before any beam is applied.
        """
        ox = np.ascontiguousarray(np.asarray(ox, dtype=np.float64).reshape(1, -1))
        oy = np.ascontiguousarray(np.asarray(oy, dtype=np.float64).reshape(1, -1))
        cx = np.asarray(center_x, dtype=np.float64).reshape(-1)[:1].copy()
        cy = np.asarray(center_y, dtype=np.float64).reshape(-1)[:1].copy()
        if ox.shape != oy.shape:
            raise ValueError("ox and oy differ in length: %s vs %s" % (ox.shape, oy.shape))
        # incremental read-back: only the 64 x 64-cell tiles this scan touched cross PCIe and are patched
        # into the host maps, so the call costs the same on a 4096^2 map as on the reference's 200^2
        count = ctypes.c_int(-1)
        rc = self._L.b2s_mapping_update_incremental_f64(
            self._h, _lib.ptr(ox), _lib.ptr(oy), _lib.ptr(cx), _lib.ptr(cy), 1, ox.shape[1], _lib.ptr(self._pmap8),
            _lib.ptr(self._tiles), self._tiles.shape[0], ctypes.byref(count))
        self._raise_nonfinite(rc)
        _lib.check(rc)
        self._host_map_blank = False         # after a reset the C side rewrites the whole host map
        if self._pmap64 is None:
            self._pmap64 = self._pmap8.astype(np.float64)
        elif count.value < 0:
            self._pmap64[...] = self._pmap8          # whole map rewritten: refresh the live array in place
        else:
            tiles_y = (self.yw + 63) // 64
            for t in self._tiles[:count.value]:
                x0, y0 = (int(t) // tiles_y) * 64, (int(t) % tiles_y) * 64
                self._pmap64[x0:x0 + 64, y0:y0 + 64] = self._pmap8[x0:x0 + 64, y0:y0 + 64]
        return self._pmap64

    def _raise_nonfinite(self, rc):
        if rc == _lib.ERR_NONFINITE:
            if "NaN" in self._L.b2s_last_error().decode():
                raise ValueError("cannot convert float NaN to integer")
            raise OverflowError("cannot convert float infinity to integer")

    # ------------------------------------------------------------------ batched entry points

    def update_batch(self, ox, oy, cx, cy, want_pmap=True):
        """K scans at once: ox, oy (K,N); cx, cy (K,).

        The coordinates are consumed in the dtype they arrive in: float32 arrays (all four) take the
        float32 entry point -- half the bytes over PCIe, exact because float32 -> float64 is -- and anything
        else is carried as float64, which is what [MAP]:33-36 evaluates int(10*(v+10)) on; the result is
        then bit-identical to K reference update() calls on the same arrays.

        Returns the refreshed occupancy as int8 (xw, yw) in {0, 50, 100} (a view of the object's
        host buffer, overwritten by the next call), or None with want_pmap=False."""
        f32 = all(getattr(a, "dtype", None) == np.float32 for a in (ox, oy, cx, cy))
        dt = np.float32 if f32 else np.float64
        ox = np.ascontiguousarray(ox, dtype=dt)
        oy = np.ascontiguousarray(oy, dtype=dt)
        cx = np.ascontiguousarray(cx, dtype=dt).reshape(-1)
        cy = np.ascontiguousarray(cy, dtype=dt).reshape(-1)
        if ox.ndim != 2 or ox.shape != oy.shape or cx.shape[0] != ox.shape[0] \
                or cy.shape[0] != ox.shape[0]:
            raise ValueError("expected ox, oy (K,N) and cx, cy (K,), got %s %s %s %s"
                             % (ox.shape, oy.shape, cx.shape, cy.shape))
        out = self._pmap8 if want_pmap else None
        self._pmap64 = None
        call = self._L.b2s_mapping_update if f32 else self._L.b2s_mapping_update_f64
        rc = call(self._h, _lib.ptr(ox), _lib.ptr(oy), _lib.ptr(cx), _lib.ptr(cy), ox.shape[0], ox.shape[1],
                  _lib.ptr(out))
        self._raise_nonfinite(rc)
        _lib.check(rc)
        if want_pmap:
            self._host_map_blank = False
        return self._pmap8 if want_pmap else None

    def update_scans(self, ranges, poses, angle_min, angle_max, clamp_inf_to=30.0, want_pmap=True):
        """Fused ingestion: raw scans + poses instead of world-frame endpoints.

        Does what slam_ekf.py:89-90 does around Mapping.update -- laserToNumpy (inf -> 30 m,
        slam_ekf.py:115-123), obs = u2T(pose).dot(np_msg) (slam_ekf.py:130-137), update(obs[0], obs[1],
        x, y) -- in one kernel, in float64, for K scans: ranges (K,N) float32, poses (K,3) = x, y, yaw.
        Half the bytes of the endpoint form cross PCIe.  Returns the int8 occupancy like update_batch.
        """
        from b2slam import scan
        ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        if ranges.ndim == 1:
            ranges = ranges.reshape(1, -1)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        if poses.shape[0] != ranges.shape[0]:
            raise ValueError("need one pose per scan, got %d poses for %d scans" % (poses.shape[0], ranges.shape[0]))
        key = (float(angle_min), float(angle_max), ranges.shape[1])
        if getattr(self, "_beam_key", None) != key:
            self._beam_cs = scan.beam_table(angle_min, angle_max, ranges.shape[1])
            self._beam_key = key
        out = self._pmap8 if want_pmap else None
        self._pmap64 = None
        # cos / sin of the yaw (u2T, slam_ekf.py:130-137) are taken inside the call, chunk by chunk, while the
        # previous chunk is on its way to the device
        rc = self._L.b2s_mapping_update_scans(self._h, _lib.ptr(ranges), _lib.ptr(poses), _lib.ptr(self._beam_cs),
                                              float(clamp_inf_to or 0.0), ranges.shape[0], ranges.shape[1],
                                              _lib.ptr(out))
        self._raise_nonfinite(rc)
        _lib.check(rc)
        if want_pmap:
            self._host_map_blank = False
        return self._pmap8 if want_pmap else None

    # ------------------------------------------------------------------ streaming calls (two steps in flight)

    def _stream_slot(self):
        if getattr(self, "_stream_maps", None) is None:
            self._stream_maps = [_lib.pinned_empty((self.xw, self.yw), np.int8) for _ in range(2)]
            self._stream_keep = [None, None]
        return self._stream_maps

    def submit_scans(self, ranges, poses, angle_min, angle_max, clamp_inf_to=30.0, zero_first=False):
        """update_scans without the wait (b2s_mapping_submit_scans): the step is enqueued -- upload, ray-cast, finalize,
        read-back of the map on a third stream -- and a MapTicket is returned at once.  Up to two steps are in flight, so
        the upload and ray-cast of the next step run while this step's 16.8 MB map (at 4096^2) crosses PCIe the other
        way.  ticket.wait() returns this step's int8 occupancy (a page-locked buffer that the second-next submit
        overwrites) or raises what update_scans would have raised for it.  zero_first: clear the counts before the step.
        The arrays must stay untouched until the ticket has been waited for."""
        from b2slam import scan
        ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        if ranges.ndim == 1:
            ranges = ranges.reshape(1, -1)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        if poses.shape[0] != ranges.shape[0]:
            raise ValueError("need one pose per scan, got %d poses for %d scans" % (poses.shape[0], ranges.shape[0]))
        key = (float(angle_min), float(angle_max), ranges.shape[1])
        if getattr(self, "_beam_key", None) != key:
            self._beam_cs = scan.beam_table(angle_min, angle_max, ranges.shape[1])
            self._beam_key = key
        maps = self._stream_slot()
        t = ctypes.c_int(-1)
        n = getattr(self, "_submitted", 0)
        out = maps[n % 2]
        self._pmap64 = None
        _lib.check(self._L.b2s_mapping_submit_scans(self._h, _lib.ptr(ranges), _lib.ptr(poses), _lib.ptr(self._beam_cs),
                                                    float(clamp_inf_to or 0.0), ranges.shape[0], ranges.shape[1],
                                                    1 if zero_first else 0, _lib.ptr(out), ctypes.byref(t)))
        self._submitted = n + 1
        self._stream_keep[n % 2] = (ranges, poses)       # keep the inputs alive while the device reads them
        return MapTicket(self, t.value, out)

    def submit_batch(self, ox, oy, cx, cy, zero_first=False):
        """update_batch (float32 endpoints) without the wait; see submit_scans."""
        ox = np.ascontiguousarray(ox, dtype=np.float32)
        oy = np.ascontiguousarray(oy, dtype=np.float32)
        cx = np.ascontiguousarray(cx, dtype=np.float32).reshape(-1)
        cy = np.ascontiguousarray(cy, dtype=np.float32).reshape(-1)
        if ox.ndim != 2 or ox.shape != oy.shape or cx.shape[0] != ox.shape[0] or cy.shape[0] != ox.shape[0]:
            raise ValueError("expected ox, oy (K,N) and cx, cy (K,), got %s %s %s %s" % (ox.shape, oy.shape, cx.shape, cy.shape))
        maps = self._stream_slot()
        t = ctypes.c_int(-1)
        n = getattr(self, "_submitted", 0)
        out = maps[n % 2]
        self._pmap64 = None
        _lib.check(self._L.b2s_mapping_submit(self._h, _lib.ptr(ox), _lib.ptr(oy), _lib.ptr(cx), _lib.ptr(cy), ox.shape[0],
                                              ox.shape[1], 1 if zero_first else 0, _lib.ptr(out), ctypes.byref(t)))
        self._submitted = n + 1
        self._stream_keep[n % 2] = (ox, oy, cx, cy)
        return MapTicket(self, t.value, out)

    def counts(self):
        """(hit, miss) int32 (xw, yw) snapshots of the device planes."""
        hit = np.empty((self.xw, self.yw), dtype=np.int32)
        miss = np.empty((self.xw, self.yw), dtype=np.int32)
        _lib.check(self._L.b2s_mapping_read(self._h, _lib.ptr(hit), _lib.ptr(miss), None, None))
        return hit, miss

    @property
    def datamap(self):
        """Evidence score [MAP]:15 as float64, = miss_weight*miss + hit_weight*hit."""
        d = np.empty((self.xw, self.yw), dtype=np.float32)
        _lib.check(self._L.b2s_mapping_read(self._h, None, None, _lib.ptr(d), None))
        return d.astype(np.float64)

    def occupancy(self):
        """int8 (xw, yw) occupancy in {0, 50, 100} straight from the device."""
        p = np.empty((self.xw, self.yw), dtype=np.int8)
        _lib.check(self._L.b2s_mapping_read(self._h, None, None, None, _lib.ptr(p)))
        return p

    # ------------------------------------------------------------------ checkpoint

    def save(self, path):
        """Write the int32 count planes and the map parameters to an .npz -- the whole state of the
        object (the reference keeps its map only in process memory)."""
        hit, miss = self.counts()
        np.savez_compressed(path, hit=hit, miss=miss, xw=self.xw, yw=self.yw, xyreso=self.xyreso,
                            hit_weight=self.hit_weight, miss_weight=self.miss_weight,
                            occ_threshold=self.occ_threshold)

    @classmethod
    def load(cls, path, device=-1):
        """Rebuild a Mapping from save(): same parameters, count planes restored bit for bit."""
        z = np.load(path)
        m = cls(int(z["xw"]), int(z["yw"]), float(z["xyreso"]), float(z["hit_weight"]), float(z["miss_weight"]),
                float(z["occ_threshold"]), device=device)
        m.set_counts(z["hit"], z["miss"])
        return m

    def set_counts(self, hit, miss):
        """Overwrite the device planes with int32 (xw, yw) host arrays (checkpoint restore)."""
        hit = np.ascontiguousarray(hit, dtype=np.int32).reshape(self.xw, self.yw)
        miss = np.ascontiguousarray(miss, dtype=np.int32).reshape(self.xw, self.yw)
        _lib.check(self._L.b2s_mapping_write(self._h, _lib.ptr(hit), _lib.ptr(miss)))
        _lib.check(self._L.b2s_mapping_read(self._h, None, None, None, _lib.ptr(self._pmap8)))
        self._host_map_blank = False
        self._pmap64 = None

    def reset(self):
        _lib.check(self._L.b2s_mapping_reset(self._h))
        self._host_map_blank = True      # the 16 MB host copy is re-filled only if somebody looks at it
        self._pmap64 = None

    def device_planes(self):
        """(hit_ptr, miss_ptr, stream_ptr) integers for layer-1 calls and collectives."""
        h, m, s = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        self._pmap64 = None      # the planes may change behind our back: the next update() rebuilds the live array
        _lib.check(self._L.b2s_mapping_planes(self._h, ctypes.byref(h), ctypes.byref(m),
                                              ctypes.byref(s)))
        return h.value, m.value, s.value or 0
