// fm_learn_vb_cuda.h -- the binding a maintainer of the reference adds next to src/libfm/src/fm_learn_vb_simultaneous.h:
// `-method vb_cuda` = the reference's own fm_learn_vb (its init(), its public fields, its files and stdout lines) with the
// loop body of fm_learn_vb_simultaneous::_learn (fm_learn_vb_simultaneous.h:18-259: update_all + test prediction +
// evaluation) delegated to the B200 engine through the C-ABI (include/svbfm.h). Compiled INSIDE the reference tree by
// oracle/make_ref_cuda.py (the only other change there: one include and one factory branch in libfm.cpp, see that script);
// tests/test_gpu_ref_tree_binding.py runs `-method vb` and `-method vb_cuda` from that one binary and compares the files.
#ifndef FM_LEARN_VB_CUDA_H_
#define FM_LEARN_VB_CUDA_H_

#include <string.h>
#include <vector>
#include "svbfm.h"
#include "fm_learn_vb.h"

// DataSubset::data_t (LargeSparseMatrix<DATA_FLOAT>, one sparse_row per attribute, fmatrix.h:36-64) -> colptr / case_id / x
static inline void svbfm_flatten_data_t(DataSubset& d, std::vector<uint64_t>& ptr, std::vector<uint32_t>& id, std::vector<float>& x) {
	ptr.assign(1, 0); id.clear(); x.clear();
	id.reserve(d.data_t->getNumValues()); x.reserve(d.data_t->getNumValues());
	for (d.data_t->begin(); !d.data_t->end(); d.data_t->next()) {
		sparse_row<DATA_FLOAT>& r = d.data_t->getRow();
		for (uint k = 0; k < r.size; k++) { id.push_back(r.data[k].id); x.push_back(r.data[k].value); }
		ptr.push_back(id.size());
	}
}

class fm_learn_vb_cuda : public fm_learn_vb {
	public:
		svbfm_t* h;
		fm_learn_vb_cuda() : h(NULL) {}
		virtual ~fm_learn_vb_cuda() { if (h) svbfm_destroy(h); }

		// init(): fm_learn_vb::init (fm_learn_vb.h:685-743) unchanged -- the variational state is drawn by the reference's own
		// code on the reference's own libc stream, so `-method vb` and `-method vb_cuda` start from the same state.

		void ck(int rc) { if (rc != 0) throw std::string(svbfm_last_error(h)); }     // -> "ERROR: ..." in main (libfm.cpp:521-525)

		virtual void learn(DataSubset& train, DataSubset& test) {
			if (task != TASK_REGRESSION) throw "vb_cuda: regression only";
			svbfm_config c; memset(&c, 0, sizeof(c));
			c.struct_size = sizeof(c); c.method = SVBFM_VB; c.num_attribute = fm->num_attribute; c.num_factor = fm->num_factor;
			c.k0 = fm->k0; c.k1 = fm->k1; c.task = 0; c.min_target = min_target; c.max_target = max_target;
			if (svbfm_create(&h, &c) != 0) throw std::string(svbfm_last_error(NULL));
			ck(svbfm_set_groups(h, meta->attr_group.value, meta->num_attr_groups));
			{
				std::vector<uint64_t> p; std::vector<uint32_t> i; std::vector<float> x;
				svbfm_flatten_data_t(train, p, i, x);
				ck(svbfm_set_csc(h, SVBFM_TRAIN, train.num_cases, (uint32_t)(p.size() - 1), &p[0], i.empty() ? NULL : &i[0], x.empty() ? NULL : &x[0], train.target.value));
				svbfm_flatten_data_t(test, p, i, x);
				ck(svbfm_set_csc(h, SVBFM_TEST, test.num_cases, (uint32_t)(p.size() - 1), &p[0], i.empty() ? NULL : &i[0], x.empty() ? NULL : &x[0], test.target.value));
			}
			// DMatrix keeps value[0] contiguous, row-major [num_factor][num_attribute] (matrix.h:104-109)
			ck(svbfm_set_state(h, mu_0_dash, sigma_0_dash, mu_w_dash.value, sigma_w_dash.value,
			                   fm->num_factor > 0 ? mu_v_dash.value[0] : NULL, fm->num_factor > 0 ? sigma_v_dash.value[0] : NULL));
			ck(svbfm_set_hyper(h, alpha, sigma_0, sigma_w.value, fm->num_factor > 0 ? sigma_v.value[0] : NULL));
			ck(svbfm_begin(h));
			std::cout << "in learn of fm_learn_vb_cuda" << std::endl;

			std::stringstream convert;
			convert << fm->k0 << fm->k1 << fm->num_factor;
			const std::string f_rmse = "test_rmse_" + convert.str() + "_vb", f_fe = "free_energy_" + convert.str() + "_vb";
			{ std::ofstream a(f_rmse.c_str()); std::ofstream b(f_fe.c_str()); }           // truncated like vbs.h:66-73
			for (uint i = 0; i < num_iter; i++) {
				double iteration_time = getusertime();
				svbfm_iter_stats s;
				ck(svbfm_vb_sweep(h, &s));
				alpha = s.alpha;
				if (s.nan_inf_count > 0) std::cout << "#nans/infs reverted:\t" << s.nan_inf_count << std::endl;
				if (s.has_free_energy) {
					std::ofstream myfile(f_fe.c_str(), std::ios_base::app);
					myfile << -s.free_energy << "\n";                                        // vb.h:678 (the file holds -F)
					std::cout << "free energy " << s.free_energy << std::endl;             // vb.h:680
				}
				std::ofstream file_rmse(f_rmse.c_str(), std::ios_base::app);
				file_rmse << s.test_rmse << "\n";                                            // vbs.h:221
				std::cout << "#Iter=" << std::setw(3) << i << "\tTrain=" << s.train_stat << "\tTest=" << s.test_rmse << std::endl;   // vbs.h:222
				if (log != NULL) {
					log->log("time_learn", getusertime() - iteration_time);
					log->log("alpha", s.alpha);
					log->log("rmse_mcmc_this", s.test_rmse);
					log->newLine();
				}
			}
			// hand the posterior back to the reference's own members (what a caller of the class can look at afterwards)
			ck(svbfm_get_state(h, &mu_0_dash, &sigma_0_dash, mu_w_dash.value, sigma_w_dash.value,
			                   fm->num_factor > 0 ? mu_v_dash.value[0] : NULL, fm->num_factor > 0 ? sigma_v_dash.value[0] : NULL));
			ck(svbfm_get_hyper(h, &alpha, &sigma_0, sigma_w.value, fm->num_factor > 0 ? sigma_v.value[0] : NULL));
		}

		// -out (libfm.cpp:514-519): the reference's body is empty (vb.h:321-348); here the clamped prediction of the final means
		virtual void predict(DataSubset& data, DVector<double>& out) {
			assert(data.num_cases == out.dim);
			ck(svbfm_predict(h, SVBFM_TEST, out.value));
		}
};

#endif /*FM_LEARN_VB_CUDA_H_*/
