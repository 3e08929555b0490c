"""Set-up utilities around the hot loop, same import paths as fbpic/lpa_utils: `boosted_frame`
(BoostConverter), `laser` (profiles, add_laser_pulse, LaserAntenna), `mirrors`, `external_fields`."""
