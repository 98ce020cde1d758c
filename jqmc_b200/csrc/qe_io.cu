// Native input path: build the engine straight from jQMC's HDF5 files, without the Python stack.
//   qe_create_from_hdf5   hamiltonian_data.h5 (dataclass tree at the root) or restart.h5 (group "hamiltonian_data")
//                         -> qe_system_desc -> qe_create         (layout: jqmc/hamiltonians.py:369-573)
//   qe_hdf5_summary       the same parse without touching CUDA (counts + checksums; CPU tests, tooling)
//   qe_hdf5_read_walkers  walker state of one rank of a restart checkpoint (jqmc/_checkpoint.py:1-30, 146-240):
//                         latest_r_up_carts / latest_r_dn_carts / jax_PRNG_key_list
// The HDF5 subset reader is qe_hdf5.hpp (this image has no libhdf5).  TREXIO files themselves are converted to
// hamiltonian_data.h5 by `jqmc-tool trexio convert-to` on the reference side; the arithmetic of that conversion
// (jqmc/trexio_wrapper.py:371-523) is mirrored in jqmc_b200/trexio_lite.py.
#include <memory>

#include "qe_common.cuh"
#include "qe_hdf5.hpp"

namespace {

struct BasisOwned {
  int cartesian = 0, n_ao = 0, n_prim = 0, n_mo = 0;
  std::vector<int32_t> nucleus_index, l, m, px, py, pz, orbital_indices;
  std::vector<double> exponents, coefficients, C;
  qe_basis_desc desc() const {
    qe_basis_desc d{};
    d.cartesian = cartesian;
    d.n_ao = n_ao;
    d.n_prim = n_prim;
    d.nucleus_index = nucleus_index.data();
    d.angular_momentums = l.data();
    d.magnetic_quantum_numbers = m.empty() ? nullptr : m.data();
    d.polynominal_order_x = px.empty() ? nullptr : px.data();
    d.polynominal_order_y = py.empty() ? nullptr : py.data();
    d.polynominal_order_z = pz.empty() ? nullptr : pz.data();
    d.orbital_indices = orbital_indices.data();
    d.exponents = exponents.data();
    d.coefficients = coefficients.data();
    d.n_mo = n_mo;
    d.mo_coefficients = C.empty() ? nullptr : C.data();
    return d;
  }
};

struct SystemOwned {
  std::vector<double> positions, zeff, lambda, j1_core, j1_Z, j_matrix, ecp_z, ecp_c;
  std::vector<int32_t> ecp_nuc, ecp_l, ecp_p, ecp_lmax;
  BasisOwned up, dn, j3;
  qe_system_desc d{};
};

double attr_num(const qeio::H5File& f, const std::string& g, const std::string& name, double dflt, bool required = false) {
  auto a = f.attrs(g);
  auto it = a.find(name);
  if (it == a.end() || it->second.cls == 3) {
    if (f.has(g + "/" + name)) return f.read(g + "/" + name).get_double(0);  // some writers store scalars as datasets
    if (required) throw std::runtime_error("hdf5: missing attribute " + g + "/" + name);
    return dflt;
  }
  return it->second.get_double(0);
}

std::string class_of(const qeio::H5File& f, const std::string& g) { return f.attr_string(g, "_class_name"); }

void read_aos(const qeio::H5File& f, const std::string& g, BasisOwned& b) {
  const std::string cls = class_of(f, g);
  if (cls != "AOs_sphe_data" && cls != "AOs_cart_data") throw std::runtime_error("hdf5: " + g + " is not an AO data group (" + cls + ")");
  b.cartesian = cls == "AOs_cart_data";
  b.n_ao = (int)attr_num(f, g, "num_ao", 0, true);
  b.n_prim = (int)attr_num(f, g, "num_ao_prim", 0, true);
  b.nucleus_index = f.read(g + "/nucleus_index").as_int();
  b.l = f.read(g + "/angular_momentums").as_int();
  b.orbital_indices = f.read(g + "/orbital_indices").as_int();
  b.exponents = f.read(g + "/exponents").as_double();
  b.coefficients = f.read(g + "/coefficients").as_double();
  if (b.cartesian) {
    b.px = f.read(g + "/polynominal_order_x").as_int();
    b.py = f.read(g + "/polynominal_order_y").as_int();
    b.pz = f.read(g + "/polynominal_order_z").as_int();
  } else {
    b.m = f.read(g + "/magnetic_quantum_numbers").as_int();
  }
  if ((int)b.nucleus_index.size() != b.n_ao || (int)b.l.size() != b.n_ao || (int)b.exponents.size() != b.n_prim ||
      (int)b.coefficients.size() != b.n_prim || (int)b.orbital_indices.size() != b.n_prim)
    throw std::runtime_error("hdf5: inconsistent AO table sizes in " + g);
}

void read_orbitals(const qeio::H5File& f, const std::string& g, BasisOwned& b) {
  if (class_of(f, g) == "MOs_data") {
    read_aos(f, g + "/aos_data", b);
    b.n_mo = (int)attr_num(f, g, "num_mo", 0, true);
    b.C = f.read(g + "/mo_coefficients").as_double();
    if ((long long)b.C.size() != (long long)b.n_mo * b.n_ao) throw std::runtime_error("hdf5: mo_coefficients shape in " + g);
  } else {
    read_aos(f, g, b);
    b.n_mo = 0;
  }
}

std::unique_ptr<SystemOwned> read_system(const qeio::H5File& f, std::string root, int Nv, int NN, int precision) {
  if (!root.empty() && root.back() == '/') root.pop_back();
  if (class_of(f, root.empty() ? "/" : root) != "Hamiltonian_data") {
    if (f.has(root + "/hamiltonian_data")) root += "/hamiltonian_data";  // restart.h5 keeps the tree one level down
    else throw std::runtime_error("hdf5: no Hamiltonian_data tree at '" + root + "'");
  }
  auto S = std::make_unique<SystemOwned>();
  const std::string st = root + "/structure_data", cp = root + "/coulomb_potential_data", wf = root + "/wavefunction_data";
  if (attr_num(f, st, "pbc_flag", 0) != 0) throw std::runtime_error("periodic systems are not supported");
  S->positions = f.read(st + "/positions").as_double();
  std::vector<double> Z = f.read(st + "/atomic_numbers").as_double();
  const int n_atom = (int)Z.size();
  if ((int)S->positions.size() != 3 * n_atom) throw std::runtime_error("hdf5: positions / atomic_numbers mismatch");
  qe_system_desc& d = S->d;
  d.n_atom = n_atom;
  d.ecp_flag = attr_num(f, cp, "ecp_flag", 0) != 0;
  S->zeff = Z;
  if (d.ecp_flag) {
    std::vector<double> zc = f.read(cp + "/z_cores").as_double();
    for (int a = 0; a < n_atom; ++a) S->zeff[a] -= zc[a];
    S->ecp_lmax = f.read(cp + "/max_ang_mom_plus_1").as_int();
    S->ecp_l = f.read(cp + "/ang_moms").as_int();
    S->ecp_nuc = f.read(cp + "/nucleus_index").as_int();
    S->ecp_z = f.read(cp + "/exponents").as_double();
    S->ecp_c = f.read(cp + "/coefficients").as_double();
    S->ecp_p = f.read(cp + "/powers").as_int();
    d.n_ecp = (int)S->ecp_l.size();
    d.ecp_nucleus_index = S->ecp_nuc.data();
    d.ecp_ang_moms = S->ecp_l.data();
    d.ecp_exponents = S->ecp_z.data();
    d.ecp_coefficients = S->ecp_c.data();
    d.ecp_powers = S->ecp_p.data();
    d.ecp_max_ang_mom_plus_1 = S->ecp_lmax.data();
  }
  d.positions = S->positions.data();
  d.effective_charges = S->zeff.data();
  const std::string gem = wf + "/geminal_data", jas = wf + "/jastrow_data";
  d.n_up = (int)attr_num(f, gem, "num_electron_up", 0, true);
  d.n_dn = (int)attr_num(f, gem, "num_electron_dn", 0, true);
  read_orbitals(f, gem + "/orb_data_up_spin", S->up);
  read_orbitals(f, gem + "/orb_data_dn_spin", S->dn);
  S->lambda = f.read(gem + "/lambda_matrix").as_double();
  d.orb_up = S->up.desc();
  d.orb_dn = S->dn.desc();
  d.lambda_matrix = S->lambda.data();
  if (f.has(jas + "/jastrow_nn_data")) throw std::runtime_error("NN Jastrow is out of scope of the walker engine");
  if (f.has(jas + "/jastrow_one_body_data")) {
    const std::string g = jas + "/jastrow_one_body_data";
    const std::string t = f.attr_string(g, "jastrow_1b_type");
    d.j1_type = t == "pade" ? 2 : 1;
    d.j1_param = attr_num(f, g, "jastrow_1b_param", 1.0, true);
    S->j1_core = f.read(g + "/core_electrons").as_double();
    S->j1_Z = f.read(g + "/structure_data/atomic_numbers").as_double();
    d.j1_core_electrons = S->j1_core.data();
    d.j1_atomic_numbers = S->j1_Z.data();
  }
  if (f.has(jas + "/jastrow_two_body_data")) {
    const std::string g = jas + "/jastrow_two_body_data";
    d.j2_type = f.attr_string(g, "jastrow_2b_type") == "exp" ? 2 : 1;
    d.j2_param = attr_num(f, g, "jastrow_2b_param", 1.0, true);
  }
  if (f.has(jas + "/jastrow_three_body_data")) {
    const std::string g = jas + "/jastrow_three_body_data";
    d.j3_flag = 1;
    read_orbitals(f, g + "/orb_data", S->j3);
    S->j_matrix = f.read(g + "/j_matrix").as_double();
    d.j3_orb = S->j3.desc();
    d.j_matrix = S->j_matrix.data();
  }
  d.Nv = Nv;
  d.NN = NN;
  d.precision = precision;
  return S;
}

double checksum(const std::vector<double>& v) {
  double s = 0;
  for (size_t i = 0; i < v.size(); ++i) s += v[i] * (double)(1 + i % 7);
  return s;
}

}  // namespace

extern "C" int qe_create_from_hdf5(const char* path, const char* group, int Nv, int NN, int precision, qe_engine** out) {
  if (!path || !out) return fail(QE_ERR_INVALID, "qe_create_from_hdf5: bad argument");
  try {
    qeio::H5File f(path);
    auto S = read_system(f, group ? group : "", Nv, NN, precision);
    return qe_create(&S->d, out);  // tables are copied to the device inside: S may go out of scope afterwards
  } catch (const std::exception& e) {
    return fail(QE_ERR_INVALID, std::string("qe_create_from_hdf5: ") + e.what());
  }
}

// counts16 = {n_atom, n_up, n_dn, n_ao, n_prim, n_mo, cartesian, n_ecp, ecp_flag, j1_type, j2_type, j3_flag, n_ao_j3, n_mo_j3, 0, 0}
// checks8  = weighted sums of {positions, effective charges, exponents, coefficients, mo_coefficients, lambda, ECP exponents, j_matrix}
extern "C" int qe_hdf5_summary(const char* path, const char* group, int64_t* counts16, double* checks8) {
  if (!path || !counts16 || !checks8) return fail(QE_ERR_INVALID, "qe_hdf5_summary: bad argument");
  try {
    qeio::H5File f(path);
    auto S = read_system(f, group ? group : "", 6, 1, 0);
    const qe_system_desc& d = S->d;
    const int64_t c[16] = {d.n_atom, d.n_up, d.n_dn, d.orb_up.n_ao, d.orb_up.n_prim, d.orb_up.n_mo, d.orb_up.cartesian, d.n_ecp,
                           d.ecp_flag, d.j1_type, d.j2_type, d.j3_flag, d.j3_flag ? d.j3_orb.n_ao : 0, d.j3_flag ? d.j3_orb.n_mo : 0, 0, 0};
    for (int i = 0; i < 16; ++i) counts16[i] = c[i];
    checks8[0] = checksum(S->positions);
    checks8[1] = checksum(S->zeff);
    checks8[2] = checksum(S->up.exponents);
    checks8[3] = checksum(S->up.coefficients);
    checks8[4] = checksum(S->up.C);
    checks8[5] = checksum(S->lambda);
    checks8[6] = checksum(S->ecp_z);
    checks8[7] = checksum(S->j_matrix);
    return QE_OK;
  } catch (const std::exception& e) {
    return fail(QE_ERR_INVALID, std::string("qe_hdf5_summary: ") + e.what());
  }
}

extern "C" int qe_hdf5_read_walkers(const char* path, int rank, int capacity_walkers, int n_up, int n_dn, int* nw, double* r_up_host,
                                    double* r_dn_host, uint32_t* keys_host) {
  if (!path || !nw || rank < 0 || n_up <= 0 || n_dn < 0) return fail(QE_ERR_INVALID, "qe_hdf5_read_walkers: bad argument");
  try {
    qeio::H5File f(path);
    const std::string g = "rank_" + std::to_string(rank);
    qeio::H5Value up = f.read(g + "/walker_state/latest_r_up_carts");
    if (up.shape.size() != 3 || (int)up.shape[1] != n_up || up.shape[2] != 3) throw std::runtime_error("latest_r_up_carts has the wrong shape");
    *nw = (int)up.shape[0];
    if (!r_up_host) return QE_OK;  // size query
    if (*nw > capacity_walkers) throw std::runtime_error("walker buffers too small");
    const std::vector<double> u = up.as_double();
    std::memcpy(r_up_host, u.data(), u.size() * sizeof(double));
    if (n_dn > 0 && r_dn_host) {
      const std::vector<double> dn = f.read(g + "/walker_state/latest_r_dn_carts").as_double();
      if (dn.size() != (size_t)*nw * n_dn * 3) throw std::runtime_error("latest_r_dn_carts has the wrong shape");
      std::memcpy(r_dn_host, dn.data(), dn.size() * sizeof(double));
    }
    if (keys_host) {
      qeio::H5Value k = f.read(g + "/rng_state/jax_PRNG_key_list");
      if (k.count() != (size_t)*nw * 2) throw std::runtime_error("jax_PRNG_key_list has the wrong shape");
      for (size_t i = 0; i < k.count(); ++i) keys_host[i] = (uint32_t)k.get_double(i);
    }
    return QE_OK;
  } catch (const std::exception& e) {
    return fail(QE_ERR_INVALID, std::string("qe_hdf5_read_walkers: ") + e.what());
  }
}
