// The pybind11 binding a maintainer of tohtsky/myFM would add next to cpp_source/declare_module.hpp
// to run create_train_fm on the CUDA engine: the reference's entry point (declare_module.hpp:30-45,
// bound at :393-394) over the C ABI of include/myfm_b200.h, with numpy / scipy arguments taken as
// py::array_t (no Eigen casters: Eigen is not needed on this side of the boundary).
//
//   _myfm_pybind.create_train_fm(rank, init_std, X, relations, y, random_seed, config, callback)
//       -> (Predictor, LearningHistory)
//   _myfm_pybind.predict_score(w0, w, V, X, relations) -> ndarray          (FM::predict_score, FM.hpp:47-136)
//
// X: scipy.sparse.csr_matrix (anything else goes through csr_matrix(), as pybind11's Eigen caster does);
// relations: sequence of objects with .original_to_block and .data (RelationBlock); config: an
// FMLearningConfig built by ConfigBuilder; callback(i, fm, hyper, history) -> bool.  The values it
// returns are instances of the classes of myfm_b200._myfm (FM, FMHyperParameters, Predictor,
// LearningHistory: the same names, members and pickle layouts as the reference's), so code written
// against `myfm._myfm` sees the types it expects.  Compiled by myfm_b200/csrc/build.py
// (g++, links libmyfm_b200.so); tests/test_gpu_pybind.py runs it against the ctypes binding.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/myfm_b200.h"

namespace py = pybind11;
using ArrD = py::array_t<double, py::array::c_style | py::array::forcecast>;
using ArrI32 = py::array_t<int32_t, py::array::c_style | py::array::forcecast>;
using ArrI64 = py::array_t<int64_t, py::array::c_style | py::array::forcecast>;

namespace {

void check(int rc) { // the exception types pybind11 raises for the reference's C++ exceptions
  if (rc == MYFM_OK)
    return;
  const std::string msg = myfm_last_error();
  if (rc == MYFM_ERR_INVALID_ARGUMENT)
    throw py::value_error(msg);
  throw std::runtime_error(msg);
}

struct Csr { // keeps the converted arrays alive while the engine reads them
  ArrI64 indptr;
  ArrI32 indices;
  ArrD data;
  myfm_csr_t view{};
  explicit Csr(py::object X) {
    py::module_ sps = py::module_::import("scipy.sparse");
    if (!py::isinstance(X, sps.attr("csr_matrix")))
      X = sps.attr("csr_matrix")(X);
    indptr = ArrI64(X.attr("indptr")), indices = ArrI32(X.attr("indices")), data = ArrD(X.attr("data"));
    const auto shape = X.attr("shape").cast<std::pair<int64_t, int64_t>>();
    view = myfm_csr_t{shape.first, shape.second, indptr.data(), indices.data(), data.data()};
  }
};

struct Relations {
  std::vector<Csr> blocks;
  std::vector<ArrI64> maps;
  std::vector<myfm_relation_t> views;
  explicit Relations(const py::sequence &rels) {
    blocks.reserve(py::len(rels));
    for (py::handle r : rels) {
      blocks.emplace_back(r.attr("data"));
      maps.emplace_back(ArrI64(r.attr("original_to_block")));
    }
    for (size_t b = 0; b < blocks.size(); b++)
      views.push_back(myfm_relation_t{maps[b].data(), static_cast<int64_t>(maps[b].size()), blocks[b].view});
  }
};

struct Config { // FMLearningConfig -> myfm_config_t
  ArrI64 group_index;
  std::vector<int32_t> n_class;
  std::vector<ArrI64> rows;
  std::vector<const int64_t *> row_ptrs;
  std::vector<int64_t> row_lens;
  myfm_config_t view{};
  explicit Config(const py::object &c) {
    view.alpha_0 = c.attr("alpha_0").cast<double>(), view.beta_0 = c.attr("beta_0").cast<double>();
    view.gamma_0 = c.attr("gamma_0").cast<double>(), view.mu_0 = c.attr("mu_0").cast<double>();
    view.reg_0 = c.attr("reg_0").cast<double>();
    view.task_type = py::int_(c.attr("task_type")).cast<int>();
    view.nu_oprobit = c.attr("nu_oprobit").cast<double>();
    view.fit_w0 = c.attr("fit_w0").cast<bool>(), view.fit_linear = c.attr("fit_linear").cast<bool>();
    view.n_iter = c.attr("n_iter").cast<int>(), view.n_kept_samples = c.attr("n_kept_samples").cast<int>();
    view.cutpoint_scale = c.attr("cutpoint_scale").cast<double>();
    group_index = ArrI64(c.attr("group_index"));
    view.group_index = group_index.data(), view.n_group_index = group_index.size();
    for (py::handle g : c.attr("cutpoint_groups")) {
      py::tuple t = py::reinterpret_borrow<py::tuple>(g);
      n_class.push_back(t[0].cast<int32_t>());
      rows.emplace_back(ArrI64(t[1]));
    }
    for (const ArrI64 &r : rows)
      row_ptrs.push_back(r.data()), row_lens.push_back(r.size());
    view.n_cutpoint_groups = static_cast<int32_t>(n_class.size());
    view.cutpoint_n_class = n_class.data(), view.cutpoint_index = row_ptrs.data(), view.cutpoint_index_len = row_lens.data();
  }
};

myfm_engine_options_t engine_options() { // dtype / rng / device as myfm_b200.options has them
  py::object o = py::module_::import("myfm_b200.options").attr("get_options")();
  myfm_engine_options_t e{};
  e.dtype = o.attr("dtype").cast<std::string>() == "f32" ? MYFM_DTYPE_F32 : MYFM_DTYPE_F64;
  e.rng = o.attr("rng").cast<std::string>() == "philox" ? MYFM_RNG_PHILOX : MYFM_RNG_MT19937;
  e.device = o.attr("device").cast<int>();
  e.world_size = 1;
  return e;
}

struct TrainerGuard {
  myfm_trainer_t *t = nullptr;
  ~TrainerGuard() { myfm_trainer_destroy(t); }
};

py::tuple create_train_fm(int rank, double init_std, py::object X, py::sequence relations, ArrD y, int random_seed,
                          py::object config, py::function callback) {
  Csr Xc(X);
  Relations rels(relations);
  Config cfg(config);
  myfm_engine_options_t opt = engine_options();
  opt.n_rows_global = Xc.view.n_rows;
  TrainerGuard guard;
  check(myfm_trainer_create(&guard.t, &Xc.view, static_cast<int32_t>(rels.views.size()), rels.views.data(), y.data(),
                            y.size(), random_seed, &cfg.view, &opt));
  check(myfm_trainer_init_fm(guard.t, rank, init_std));
  int64_t n_train = 0, dim_all = 0;
  int32_t K = 0, G = 0;
  check(myfm_trainer_dims(guard.t, &n_train, &dim_all, &K, &G));

  py::module_ types = py::module_::import("myfm_b200._myfm");
  py::object predictor = types.attr("Predictor")(rank, dim_all, cfg.view.task_type);
  py::object history = types.attr("LearningHistory")();
  py::list samples = predictor.attr("samples"), hypers = history.attr("hypers");
  for (int it = 0; it < cfg.view.n_iter; it++) { // GibbsFMTrainer::learn_with_callback, FMTrainer.hpp:66-82
    check(myfm_trainer_step(guard.t, 1));
    double w0 = 0, alpha = 0;
    ArrD w(dim_all), V({dim_all, static_cast<int64_t>(K)});
    ArrD mu_w(G), lambda_w(G), mu_V({static_cast<int64_t>(G), static_cast<int64_t>(K)}),
        lambda_V({static_cast<int64_t>(G), static_cast<int64_t>(K)});
    check(myfm_trainer_get_fm(guard.t, &w0, w.mutable_data(), V.mutable_data()));
    check(myfm_trainer_get_hyper(guard.t, &alpha, mu_w.mutable_data(), lambda_w.mutable_data(), mu_V.mutable_data(),
                                 lambda_V.mutable_data()));
    py::list cutpoints;
    for (int g = 0; g < cfg.view.n_cutpoint_groups; g++) {
      ArrD c(cfg.n_class[g] - 1);
      check(myfm_trainer_get_cutpoints(guard.t, g, c.mutable_data()));
      cutpoints.append(c);
    }
    py::object fm = types.attr("FM")(w0, w, V, cutpoints);
    py::object hyper = types.attr("FMHyperParameters")(alpha, mu_w, lambda_w, mu_V, lambda_V);
    if (cfg.view.n_iter <= it + cfg.view.n_kept_samples)
      samples.append(fm);
    hypers.append(hyper);
    if (callback(it, fm, hyper, history).cast<bool>()) // std::function<bool(int, FM*, Hyper*, History*)>
      break;
  }
  check(myfm_trainer_sync(guard.t));
  py::list accepts = history.attr("n_mh_accept");
  for (int g = 0; g < cfg.view.n_cutpoint_groups; g++) {
    int64_t n = 0;
    check(myfm_trainer_mh_accept(guard.t, g, &n));
    accepts.append(n);
  }
  return py::make_tuple(predictor, history);
}

ArrD predict_score(double w0, ArrD w, ArrD V, py::object X, py::sequence relations) {
  Csr Xc(X);
  Relations rels(relations);
  myfm_engine_options_t opt = engine_options();
  myfm_dataset_t *d = nullptr;
  check(myfm_dataset_create(&d, &Xc.view, static_cast<int32_t>(rels.views.size()), rels.views.data(), opt.dtype,
                            opt.device));
  ArrD out(Xc.view.n_rows);
  const int rc = myfm_predict_score(d, w0, w.data(), V.data(), w.size(), V.ndim() == 2 ? static_cast<int32_t>(V.shape(1)) : 0,
                                    out.mutable_data());
  myfm_dataset_destroy(d);
  check(rc);
  return out;
}

} // namespace

PYBIND11_MODULE(_myfm_pybind, m) {
  m.doc() = "pybind11 binding of the myfm_b200 C ABI: the reference's create_train_fm on the CUDA engine";
  m.def("create_train_fm", &create_train_fm, "create and train fm.", py::arg("rank"), py::arg("init_std"), py::arg("X"),
        py::arg("relations"), py::arg("y"), py::arg("random_seed"), py::arg("config"), py::arg("callback"));
  m.def("predict_score", &predict_score, py::arg("w0"), py::arg("w"), py::arg("V"), py::arg("X"),
        py::arg("relations") = py::list());
  m.def("device_count", [] {
    int32_t n = 0;
    check(myfm_device_count(&n));
    return n;
  });
}
