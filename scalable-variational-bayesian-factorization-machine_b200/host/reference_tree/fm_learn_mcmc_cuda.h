// fm_learn_mcmc_cuda.h -- `-method mcmc_cuda` / `-method als_cuda` inside the reference tree: the reference's own
// fm_learn_mcmc (init(), public fields w_lambda / v_lambda / do_sample / do_multilevel poked by main, libfm.cpp:297-305,
// 367-405) with the loop of fm_learn_mcmc_simultaneous::_learn (fm_learn_mcmc_simultaneous.h:50-305) delegated to the
// B200 engine. Sampled runs differ from `-method mcmc` in the draws (Philox instead of libc rand()), als is deterministic.
// See fm_learn_vb_cuda.h and oracle/make_ref_cuda.py.
#ifndef FM_LEARN_MCMC_CUDA_H_
#define FM_LEARN_MCMC_CUDA_H_

#include "fm_learn_vb_cuda.h"
#include "fm_learn_mcmc.h"

class fm_learn_mcmc_cuda : public fm_learn_mcmc {
	public:
		svbfm_t* h;
		long seed;
		fm_learn_mcmc_cuda() : h(NULL), seed(0) {}
		virtual ~fm_learn_mcmc_cuda() { if (h) svbfm_destroy(h); }
		void ck(int rc) { if (rc != 0) throw std::string(svbfm_last_error(h)); }

		virtual void learn(DataSubset& train, DataSubset& test) {
			if (task != TASK_REGRESSION) throw "mcmc_cuda: regression only in this binding";
			svbfm_config c; memset(&c, 0, sizeof(c));
			c.struct_size = sizeof(c); c.method = SVBFM_MCMC; c.num_attribute = fm->num_attribute; c.num_factor = fm->num_factor;
			c.k0 = fm->k0; c.k1 = fm->k1; c.task = 0; c.min_target = min_target; c.max_target = max_target;
			c.do_sample = do_sample; c.do_multilevel = do_multilevel; c.seed = (uint64_t)seed;
			c.reg0 = fm->reg0; c.regw = fm->regw; c.regv = fm->regv;
			if (svbfm_create(&h, &c) != 0) throw std::string(svbfm_last_error(NULL));
			ck(svbfm_set_groups(h, meta->attr_group.value, meta->num_attr_groups));
			{
				std::vector<uint64_t> p; std::vector<uint32_t> i; std::vector<float> x;
				svbfm_flatten_data_t(train, p, i, x);
				ck(svbfm_set_csc(h, SVBFM_TRAIN, train.num_cases, (uint32_t)(p.size() - 1), &p[0], i.empty() ? NULL : &i[0], x.empty() ? NULL : &x[0], train.target.value));
				svbfm_flatten_data_t(test, p, i, x);
				ck(svbfm_set_csc(h, SVBFM_TEST, test.num_cases, (uint32_t)(p.size() - 1), &p[0], i.empty() ? NULL : &i[0], x.empty() ? NULL : &x[0], test.target.value));
			}
			// the parameters of mcmc live in fm_model (fm_model.h:44-46), drawn by fm.init() and fm.w.init_normal (libfm.cpp:273, 298)
			ck(svbfm_set_state(h, fm->w0, 0.0, fm->w.value, NULL, fm->num_factor > 0 ? fm->v.value[0] : NULL, NULL));
			// per-group regularisation as main() left it in w_lambda / v_lambda (libfm.cpp:367-405)
			ck(svbfm_set_hyper(h, alpha, fm->reg0, w_lambda.value, fm->num_factor > 0 ? v_lambda.value[0] : NULL));
			ck(svbfm_begin(h));

			std::stringstream convert;
			convert << fm->k0 << fm->k1 << fm->num_factor;
			const std::string f_rmse = "test_rmse_" + convert.str() + "_mcmc";
			{ std::ofstream a(f_rmse.c_str()); }                                            // mcmcs.h:60-62
			for (uint i = 0; i < num_iter; i++) {
				double iteration_time = getusertime();
				svbfm_iter_stats s;
				ck(svbfm_mcmc_sweep(h, &s));
				alpha = s.alpha;
				if (s.nan_inf_count > 0) std::cout << "#nans/infs reverted:\t" << s.nan_inf_count << std::endl;
				std::cout << "#Iter=" << std::setw(3) << i << "\tTrain=" << s.train_stat << "\tTest=" << s.test_rmse << std::endl;   // mcmcs.h:244
				std::ofstream file_rmse(f_rmse.c_str(), std::ios_base::app);
				file_rmse << s.test_rmse << "\n";                                            // mcmcs.h:245
				if (log != NULL) {
					log->log("time_learn", getusertime() - iteration_time);
					log->log("alpha", s.alpha);
					log->log("rmse", s.test_rmse);
					log->log("rmse_mcmc_this", s.rmse_this);
					log->log("rmse_mcmc_all", s.test_rmse);
					log->newLine();
				}
			}
			ck(svbfm_get_state(h, &fm->w0, NULL, fm->w.value, NULL, fm->num_factor > 0 ? fm->v.value[0] : NULL, NULL));
		}

		virtual void predict(DataSubset& data, DVector<double>& out) {                      // mcmc.h:355-379
			assert(data.num_cases == out.dim);
			ck(svbfm_predict(h, SVBFM_TEST, out.value));
		}
};

#endif /*FM_LEARN_MCMC_CUDA_H_*/
