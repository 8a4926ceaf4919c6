"""Drop-in alias: `import efa_xray...` resolves to the B200 implementation in efa_xray_b200, module for
module (efa_xray.state.ensemble, efa_xray.observation.observation, efa_xray.assimilation.ensrf,
efa_xray.assimilation.assimilation).  Nothing is implemented here."""
